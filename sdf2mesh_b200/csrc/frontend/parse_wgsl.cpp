// parse_wgsl.cpp -- WGSL front-end (the subset SDF code uses; see DESIGN.md section 5 for the
// list).  Replaces naga's WGSL front-end + validator behind wgpu's create_shader_module
// (/root/reference/src/shader.rs:220-225).
#include "parse.h"
#include "parser_base.h"

namespace s2m_frontend {

namespace {

class WgslParser : public ParserBase {
 public:
  WgslParser(Module* m, const std::vector<std::string>& builtin_fns) : ParserBase(Lang::Wgsl, m), builtin_fns_(builtin_fns.begin(), builtin_fns.end()) {}

  void parse(const std::string& src) {
    LexOptions lo;
    toks = Lexer(src, lo).run();
    push_scope();
    collect_module_scope();
    resolve_globals();      // constants, variables and struct bodies (order-independent)
    resolve_signatures();
    for (auto& fb : fn_bodies_) parse_function_body(fb);
  }

 private:
  std::set<std::string> builtin_fns_;
  struct GlobalDecl { size_t tok; bool done = false; StructDef* sdef = nullptr; };
  struct FnBody { Function* fn; size_t tok; };
  std::vector<GlobalDecl> global_decls_;
  struct FnSig { Function* fn; size_t tok; };
  std::vector<FnSig> fn_sigs_;             // token index of each non-entry function's `(`
  std::vector<FnBody> fn_bodies_;
  std::map<std::string, Type> aliases_;

  // ---------------------------------------------------------------- types
  bool is_type_name(const std::string& s) const {
    static const std::set<std::string> names = {"f32", "i32", "u32", "bool", "vec2", "vec3", "vec4", "vec2f", "vec3f", "vec4f",
                                                "vec2i", "vec3i", "vec4i", "vec2u", "vec3u", "vec4u", "vec2h", "vec3h", "vec4h", "f16",
                                                "mat2x2", "mat3x3", "mat4x4", "mat2x2f", "mat3x3f", "mat4x4f"};
    return names.count(s) || aliases_.count(s) || structs.count(s) || s == "array";
  }
  static bool scalar_kind(const std::string& s, Sk* sk) {
    if (s == "f32") { *sk = Sk::F32; return true; }
    if (s == "i32") { *sk = Sk::I32; return true; }
    if (s == "u32") { *sk = Sk::U32; return true; }
    if (s == "bool") { *sk = Sk::Bool; return true; }
    return false;
  }
  // closes a template list; splits a `>>` token in two (`ptr<function, vec4<f32>>`)
  void expect_close_angle() {
    if (is_punct(">>")) { toks[pos].text = ">"; return; }
    expect(">");
  }
  // Parses a type.  *infer is set when the scalar kind was left to inference (`vec3` without <T>).
  Type parse_type(bool* infer = nullptr, bool* is_ptr = nullptr) {
    if (infer) *infer = false;
    const std::string name = expect_ident("a type");
    Sk sk;
    if (scalar_kind(name, &sk)) return Type::scalar(sk);
    if (aliases_.count(name)) return aliases_[name];
    if (name == "f16" || (name.size() == 5 && name.compare(0, 3, "vec") == 0 && name[4] == 'h')) b.unsupported("f16 types");
    if (name.size() >= 4 && name.compare(0, 3, "vec") == 0 && name[3] >= '2' && name[3] <= '4') {
      const int n = name[3] - '0';
      if (name.size() == 5) {
        const char c = name[4];
        return Type::vec(c == 'f' ? Sk::F32 : c == 'i' ? Sk::I32 : Sk::U32, n);
      }
      if (accept("<")) {
        const std::string el = expect_ident("a scalar type");
        if (!scalar_kind(el, &sk)) perr("unknown vector component type " + el);
        expect_close_angle();
        return Type::vec(sk, n);
      }
      if (infer) { *infer = true; return Type::vec(Sk::F32, n); }
      perr("vector type needs a component type");
    }
    if (name == "ptr") {
      expect("<");
      expect_ident("an address space");
      expect(",");
      Type t = parse_type();
      if (accept(",")) expect_ident("an access mode");
      expect_close_angle();
      if (!is_ptr) b.unsupported("pointer types outside of function parameters");
      *is_ptr = true;
      return t;
    }
    if (name.compare(0, 3, "mat") == 0) {
      if (name.size() >= 6 && name[4] == 'x' && name[3] == name[5] && name[3] >= '2' && name[3] <= '4' && (name.size() == 6 || (name.size() == 7 && name[6] == 'f'))) {
        if (name.size() == 6 && accept("<")) {
          if (expect_ident("a scalar type") != "f32") b.unsupported("non-f32 matrices");
          expect_close_angle();
        }
        return Type::mat(name[3] - '0');
      }
      b.unsupported("matrix type " + name + " (only square f32 matrices)");
    }
    if (name == "array") {
      if (!accept("<")) perr("array type needs an element type (array<T, N>)");
      Type el = parse_type();
      if (!accept(",")) b.unsupported("runtime-sized arrays");
      const int len = array_len(parse_binary(9));  // additive level: `>` closes the template list
      expect_close_angle();
      return mod->array_of(el, len);
    }
    if (structs.count(name)) return Type::struct_(structs[name]);
    if (name == "atomic" || name.compare(0, 7, "texture") == 0 || name == "sampler") b.unsupported("type " + name);
    b.unsupported("user-defined type '" + name + "'");
  }

  void skip_attributes() {
    while (accept("@")) {
      expect_ident("an attribute name");
      if (is_punct("(")) {
        int depth = 0;
        do {
          if (is_punct("(")) ++depth;
          if (is_punct(")")) --depth;
          advance();
        } while (depth > 0 && peek().k != Token::End);
      }
    }
  }

  // ---------------------------------------------------------------- module scope
  void collect_module_scope() {
    while (peek().k != Token::End) {
      if (accept(";")) continue;
      const size_t start = pos;
      bool entry = false;
      while (is_punct("@")) {
        if (is_ident("fragment", 1) || is_ident("vertex", 1) || is_ident("compute", 1)) entry = true;
        size_t save = pos;
        skip_attributes();
        if (pos == save) break;
      }
      if (is_ident("fn")) { collect_function(entry); continue; }
      if (is_ident("const") || is_ident("var") || is_ident("override") || is_ident("let")) {
        global_decls_.push_back({start});
        while (!is_punct(";") && peek().k != Token::End) advance();
        expect(";");
        continue;
      }
      if (accept_ident("alias")) {
        const std::string n = expect_ident("alias name");
        expect("=");
        aliases_[n] = parse_type();
        expect(";");
        continue;
      }
      if (is_ident("struct")) {
        GlobalDecl gd{pos};
        advance();
        gd.sdef = declare_struct(expect_ident("struct name"));
        global_decls_.push_back(gd);
        skip_braces();
        continue;
      }
      if (is_ident("enable") || is_ident("requires") || is_ident("diagnostic")) {
        while (!is_punct(";") && peek().k != Token::End) advance();
        expect(";");
        continue;
      }
      if (is_ident("const_assert")) { while (!is_punct(";") && peek().k != Token::End) advance(); expect(";"); continue; }
      perr("expected a module-scope declaration");
    }
  }

  void collect_function(bool entry) {
    advance();  // fn
    const std::string name = expect_ident("function name");
    b.cur_line = peek().line;
    if (functions.count(name)) b.error("redefinition of function '" + name + "'");
    mod->functions.emplace_back(new Function());
    Function* fn = mod->functions.back().get();
    fn->name = name; fn->line = peek().line; fn->is_entry = entry;
    fn->builtin_lib = builtin_fns_.count(name) > 0 && name != "sdf3d_normal";
    functions[name] = fn;
    // types in the signature may name structs / constants declared further down: parse it later
    fn_sigs_.push_back({fn, pos});
    while (!is_punct("{") && peek().k != Token::End) advance();
    if (!is_punct("{")) perr("expected function body");
    if (!entry) fn_bodies_.push_back({fn, pos});
    skip_braces();
  }

  void resolve_signatures() {
    for (const FnSig& sg : fn_sigs_) {
      pos = sg.tok;
      Function* fn = sg.fn;
      const bool entry = fn->is_entry;
      b.cur_line = peek().line;
      expect("(");
      while (!is_punct(")")) {
        skip_attributes();
        const std::string pn = expect_ident("parameter name");
        expect(":");
        skip_attributes();
        if (entry) {  // entry-point signatures may use types we do not model; skip them
          int depth = 0;
          while (peek().k != Token::End && !(depth == 0 && (is_punct(",") || is_punct(")")))) {
            if (is_punct("<") || is_punct("(")) ++depth;
            if (is_punct(">") || is_punct(")")) --depth;
            advance();
          }
        } else {
          bool is_ptr = false;
          Type t = parse_type(nullptr, &is_ptr);
          Var* v = mod->new_var();
          v->name = pn; v->ty = t; v->storage = Var::Param; v->immutable = true;
          v->is_ptr = is_ptr; v->by_ref = is_ptr;
          fn->params.push_back(v);
        }
        if (!accept(",")) break;
      }
      expect(")");
      fn->ret = Type::void_();
      if (accept("->")) {
        skip_attributes();
        if (!entry) fn->ret = parse_type();
      }
    }
  }

  void resolve_globals() {
    // module-scope declarations are order-independent in WGSL: retry until no progress
    size_t remaining = global_decls_.size();
    std::string last_err;
    for (;;) {
      size_t progress = 0;
      for (auto& g : global_decls_) {
        if (g.done) continue;
        pos = g.tok;
        try {
          parse_global(g);
          g.done = true;
          ++progress;
        } catch (const FrontendError& e) {
          if (std::string(e.what()).find("unknown identifier") == std::string::npos) throw;
          last_err = e.what();
        }
      }
      remaining -= progress;
      if (remaining == 0) break;
      if (progress == 0) throw FrontendError(4, last_err);
    }
  }

  void parse_struct_body(StructDef* d) {  // struct Name { a: T, b: U, }
    d->field_names.clear();
    d->field_types.clear();
    advance();  // struct
    advance();  // name
    expect("{");
    while (!is_punct("}")) {
      skip_attributes();
      const std::string fname = expect_ident("a member name");
      expect(":");
      skip_attributes();
      add_field(d, fname, parse_type());
      if (!accept(",")) break;
    }
    expect("}");
    if (d->field_names.empty()) b.error("struct " + d->name + " has no members");
  }

  void parse_global(GlobalDecl& g) {
    if (g.sdef) { parse_struct_body(g.sdef); return; }
    bool resource = false;
    while (is_punct("@")) { resource = true; skip_attributes(); }
    b.cur_line = peek().line;
    if (accept_ident("override")) b.unsupported("override declarations");
    if (accept_ident("let")) perr("module-scope let is not WGSL; use const");
    const bool is_const = is_ident("const");
    advance();  // const | var
    bool uniformish = resource;
    if (!is_const && accept("<")) {
      const std::string space = expect_ident("address space");
      if (space == "uniform" || space == "storage") uniformish = true;
      else if (space != "private" && space != "workgroup") perr("unknown address space " + space);
      if (accept(",")) expect_ident("access mode");
      expect(">");
    }
    const std::string name = expect_ident("a name");
    bool has_type = false, infer = false;
    Type ty;
    if (accept(":")) { ty = parse_type(&infer); has_type = !infer; }
    ExprP init;
    if (accept("=")) init = parse_expr();
    expect(";");
    if (!has_type && !init) b.error("declaration of '" + name + "' needs a type or an initializer");
    if (uniformish) init = nullptr;  // bound resources read as zero (nothing is bound in this engine)
    if (init) {
      if (has_type) init = b.coerce(init, ty, "initializer");
      else if (!is_const) init = b.concretize(init);
      ty = init->ty;
    }
    Var* v = declare(name, ty, is_const ? Var::ModuleConst : Var::Global);
    v->immutable = is_const;
    if (init) {
      ConstVal cv;
      if (b.const_eval(*init, &cv)) { v->has_const = true; v->cval = cv; }
      else if (is_const && !(init->ty.is_matrix() || init->ty.is_aggregate()) ) b.error("initializer of const '" + name + "' is not a constant expression");
      else if (is_const && !b.is_const_expr(*init)) b.error("initializer of const '" + name + "' is not a constant expression");
    } else {
      v->has_const = true;  // zero value
      v->cval.ty = ty;
    }
    mod->globals.push_back(v);
    if (init) mod->global_init[v] = init;
  }

  // ---------------------------------------------------------------- functions
  void parse_function_body(const FnBody& fb) {
    pos = fb.tok;
    cur_fn = fb.fn;
    push_scope();
    for (Var* p : fb.fn->params) {
      if (scopes.back().count(p->name)) b.error("duplicate parameter '" + p->name + "'");
      scopes.back()[p->name] = p;
    }
    fb.fn->body = parse_block();
    pop_scope();
    cur_fn = nullptr;
  }

  StmtP parse_block() {
    expect("{");
    StmtP blk = mk_stmt(Stmt::Block);
    push_scope();
    while (!is_punct("}")) {
      if (peek().k == Token::End) perr("unterminated block");
      StmtP s = parse_statement();
      if (s) blk->body.push_back(s);
    }
    expect("}");
    pop_scope();
    return blk;
  }

  StmtP parse_var_statement() {
    const std::string kw = advance().text;  // let | var | const
    if (kw == "var" && accept("<")) { expect_ident("address space"); expect(">"); }
    const std::string name = expect_ident("a variable name");
    bool has_type = false, infer = false;
    Type ty;
    if (accept(":")) { ty = parse_type(&infer); has_type = !infer; }
    ExprP init;
    if (accept("=")) init = parse_expr();
    expect(";");
    if (kw != "var" && !init) b.error(kw + " '" + name + "' needs an initializer");
    if (!has_type && !init) b.error("var '" + name + "' needs a type or an initializer");
    if (init) {
      if (init->ty.is_void()) b.error("initializer of '" + name + "' has no value");
      if (has_type) init = b.coerce(init, ty, "initializer");
      else if (kw != "const") init = b.concretize(init);
      ty = init->ty;
    }
    Var* v = declare(name, ty, Var::Local);
    v->immutable = kw != "var";
    if (kw == "const") {
      ConstVal cv;
      if (b.const_eval(*init, &cv)) {
        v->has_const = true; v->cval = cv;
        return nullptr;  // folded at every use
      }
      if (!((init->ty.is_aggregate() || init->ty.is_matrix()) && b.is_const_expr(*init))) b.error("initializer of const '" + name + "' is not a constant expression");
      // aggregate constants are not folded: they behave like `let`
    }
    StmtP s = mk_stmt(Stmt::VarDecl);
    s->var = v; s->a = init;
    return s;
  }

  // assignment / increment / call / phony assignment, without the trailing ';'
  StmtP parse_simple_statement() {
    if (is_ident("_") && is_punct("=", 1)) { advance(); advance(); parse_expr(); return nullptr; }
    ExprP lhs = parse_unary();
    if (is_punct("=")) {
      advance();
      return make_assign(lhs, parse_expr());
    }
    if (peek().k == Token::Punct) {
      const std::string p = peek().text;
      if (p == "+=" || p == "-=" || p == "*=" || p == "/=" || p == "%=" || p == "&=" || p == "|=" || p == "^=" || p == "<<=" || p == ">>=") {
        advance();
        ExprP rhs = parse_expr();
        return make_assign(lhs, b.binary(compound_op(p), lhs, rhs));
      }
      if (p == "++" || p == "--") {
        advance();
        return make_assign(lhs, b.binary(p == "++" ? Op::Add : Op::Sub, lhs, b.lit_int(1, Sk::AInt)));
      }
    }
    if (lhs->k == Expr::UserCall || lhs->k == Expr::Call) {
      StmtP s = mk_stmt(Stmt::CallStmt);
      s->a = lhs;
      return s;
    }
    perr("expected a statement");
  }

  StmtP parse_statement() {
    b.cur_line = peek().line;
    if (accept(";")) return nullptr;
    if (is_punct("{")) return parse_block();
    if (is_ident("let") || is_ident("var") || is_ident("const")) return parse_var_statement();
    if (accept_ident("return")) {
      StmtP s = mk_stmt(Stmt::Return);
      if (!is_punct(";")) {
        if (cur_fn->ret.is_void()) b.error("return with a value in a function without a return type");
        s->a = b.coerce(parse_expr(), cur_fn->ret, "return");
      } else if (!cur_fn->ret.is_void()) b.error("return without a value");
      expect(";");
      return s;
    }
    if (accept_ident("if")) return parse_if();
    if (accept_ident("for")) {
      StmtP s = mk_stmt(Stmt::For);
      push_scope();
      expect("(");
      if (!is_punct(";")) {
        if (is_ident("let") || is_ident("var") || is_ident("const")) s->init = parse_var_statement();
        else { s->init = parse_simple_statement(); expect(";"); }
      } else expect(";");
      if (!is_punct(";")) s->a = parse_condition();
      expect(";");
      if (!is_punct(")")) s->cont = parse_simple_statement();
      expect(")");
      ++loop_depth;
      s->body.push_back(parse_block());
      --loop_depth;
      pop_scope();
      return s;
    }
    if (accept_ident("while")) {
      StmtP s = mk_stmt(Stmt::While);
      s->a = parse_condition();
      ++loop_depth;
      s->body.push_back(parse_block());
      --loop_depth;
      return s;
    }
    if (accept_ident("loop")) {
      StmtP s = mk_stmt(Stmt::Loop);
      expect("{");
      push_scope();
      ++loop_depth;
      StmtP blk = mk_stmt(Stmt::Block);
      while (!is_punct("}")) {
        if (peek().k == Token::End) perr("unterminated loop");
        if (accept_ident("continuing")) {
          expect("{");
          StmtP c = mk_stmt(Stmt::Block);
          while (!is_punct("}")) {
            if (is_ident("break") && is_ident("if", 1)) {
              advance(); advance();
              s->break_if = parse_condition();
              expect(";");
              continue;
            }
            StmtP cs = parse_statement();
            if (cs) c->body.push_back(cs);
          }
          expect("}");
          s->cont = c;
          break;
        }
        StmtP st = parse_statement();
        if (st) blk->body.push_back(st);
      }
      expect("}");
      --loop_depth;
      pop_scope();
      s->body.push_back(blk);
      return s;
    }
    if (accept_ident("break")) { if (!loop_depth && !switch_depth) b.error("break outside of a loop or switch"); expect(";"); return mk_stmt(Stmt::Break); }
    if (accept_ident("continue")) { if (!loop_depth) b.error("continue outside of a loop"); expect(";"); return mk_stmt(Stmt::Continue); }
    if (accept_ident("discard")) { expect(";"); return mk_stmt(Stmt::Discard); }
    if (accept_ident("switch")) return parse_switch();
    StmtP s = parse_simple_statement();
    expect(";");
    return s;
  }

  // switch e { case 1, 2: { } case 3 { } default: { } }   (`default` may also appear in a case list)
  StmtP parse_switch() {
    StmtP sw = mk_stmt(Stmt::Switch);
    sw->a = switch_selector(parse_expr());
    expect("{");
    ++switch_depth;
    while (!is_punct("}")) {
      if (peek().k == Token::End) perr("unterminated switch");
      StmtP c = mk_stmt(Stmt::Case);
      if (accept_ident("default")) c->is_default = true;
      else if (accept_ident("case")) {
        for (;;) {
          if (accept_ident("default")) c->is_default = true;
          else c->case_values.push_back(case_value(parse_expr(), sw->a->ty));
          if (!accept(",")) break;
          if (is_punct(":") || is_punct("{")) break;  // trailing comma
        }
      } else perr("expected 'case' or 'default'");
      accept(":");
      c->body.push_back(parse_block());
      sw->body.push_back(c);
    }
    expect("}");
    --switch_depth;
    check_cases(*sw);
    bool has_default = false;
    for (const StmtP& c : sw->body) has_default = has_default || c->is_default;
    if (!has_default) b.error("switch needs a default clause");
    return sw;
  }

  ExprP parse_condition() {
    ExprP c = parse_expr();
    if (!c->ty.is_bool() || !c->ty.is_scalar()) b.error("condition must be a bool, found " + c->ty.str());
    return c;
  }

  StmtP parse_if() {
    StmtP s = mk_stmt(Stmt::If);
    s->a = parse_condition();
    s->then_s = parse_block();
    if (accept_ident("else")) {
      if (accept_ident("if")) s->else_s = parse_if();
      else s->else_s = parse_block();
    }
    return s;
  }

  // ---------------------------------------------------------------- expressions
  ExprP parse_expr() { return parse_binary(0); }

  static int prec_of(const std::string& p) {
    if (p == "||") return 1;
    if (p == "&&") return 2;
    if (p == "|") return 3;
    if (p == "^") return 4;
    if (p == "&") return 5;
    if (p == "==" || p == "!=") return 6;
    if (p == "<" || p == ">" || p == "<=" || p == ">=") return 7;
    if (p == "<<" || p == ">>") return 8;
    if (p == "+" || p == "-") return 9;
    if (p == "*" || p == "/" || p == "%") return 10;
    return -1;
  }
  static Op op_of(const std::string& p) {
    if (p == "||") return Op::Or;
    if (p == "&&") return Op::And;
    if (p == "|") return Op::BitOr;
    if (p == "^") return Op::BitXor;
    if (p == "&") return Op::BitAnd;
    if (p == "==") return Op::Eq;
    if (p == "!=") return Op::Ne;
    if (p == "<") return Op::Lt;
    if (p == ">") return Op::Gt;
    if (p == "<=") return Op::Le;
    if (p == ">=") return Op::Ge;
    if (p == "<<") return Op::Shl;
    if (p == ">>") return Op::Shr;
    if (p == "+") return Op::Add;
    if (p == "-") return Op::Sub;
    if (p == "*") return Op::Mul;
    if (p == "/") return Op::Div;
    return Op::Rem;
  }

  ExprP parse_binary(int min_prec) {
    ExprP lhs = parse_unary();
    for (;;) {
      if (peek().k != Token::Punct) break;
      const std::string p = peek().text;
      const int prec = prec_of(p);
      if (prec < 0 || prec < min_prec) break;
      advance();
      ExprP rhs = parse_binary(prec + 1);
      lhs = b.binary(op_of(p), lhs, rhs);
    }
    return lhs;
  }

  ExprP parse_unary() {
    b.cur_line = peek().line;
    if (accept("-")) return b.unary(Op::Neg, parse_unary());
    if (accept("!")) return b.unary(Op::Not, parse_unary());
    if (accept("~")) return b.unary(Op::BitNot, parse_unary());
    if (accept("&")) return b.addr_of(parse_unary());
    if (accept("*")) return b.deref(parse_unary());
    return parse_postfix(parse_primary());
  }

  ExprP parse_postfix(ExprP e) {
    for (;;) {
      if (accept(".")) {
        const std::string m = expect_ident("a member name");
        if (e->k == Expr::VarRef && e->var->is_ptr) e = b.deref(e);  // p.x on a pointer parameter
        e = b.member(e, m);
        continue;
      }
      if (is_punct("[")) {
        advance();
        ExprP idx = parse_expr();
        expect("]");
        if (e->k == Expr::VarRef && e->var->is_ptr) e = b.deref(e);
        e = b.index(e, idx);
        continue;
      }
      break;
    }
    return e;
  }

  std::vector<ExprP> parse_args() {
    std::vector<ExprP> args;
    expect("(");
    while (!is_punct(")")) {
      args.push_back(parse_expr());
      if (!accept(",")) break;
    }
    expect(")");
    return args;
  }

  ExprP parse_primary() {
    const Token& t = peek();
    b.cur_line = t.line;
    if (t.k == Token::Float) {
      advance();
      return b.lit_float(t.fval, t.suffix == 'f' ? Sk::F32 : Sk::AFloat);
    }
    if (t.k == Token::Int) {
      advance();
      if (t.suffix == 'u') return b.lit_int(t.ival, Sk::U32);
      if (t.suffix == 'i') return b.lit_int(t.ival, Sk::I32);
      if (t.suffix == 'f') return b.lit_float((double)t.ival, Sk::F32);
      return b.lit_int(t.ival, Sk::AInt);
    }
    if (accept("(")) {
      ExprP e = parse_expr();
      expect(")");
      return e;
    }
    if (t.k != Token::Ident) perr("expected an expression");
    const std::string name = t.text;
    if (name == "true" || name == "false") { advance(); return b.lit_bool(name == "true"); }
    // a local / global variable shadows everything else
    if (Var* v = lookup(name)) {
      advance();
      if (is_punct("(")) b.error("'" + name + "' is a variable, not a function");
      return b.var_ref(v);
    }
    if (name == "array" && is_punct("(", 1)) {  // array(a, b, c): element type and length inferred
      advance();
      std::vector<ExprP> args = parse_args();
      if (args.empty()) b.error("cannot infer the type of an empty array constructor");
      bool any_float = false;
      const Expr* concrete = nullptr;
      for (const ExprP& a : args) { if (!a->ty.is_abstract() && !concrete) concrete = a.get(); if (a->ty.is_float()) any_float = true; }
      Type el = concrete ? concrete->ty : any_float ? args[0]->ty.with_sk(Sk::F32) : args[0]->ty.with_sk(Sk::I32);
      return b.construct(mod->array_of(el, (int)args.size()), false, args);
    }
    if (is_type_name(name) && (is_punct("(", 1) || is_punct("<", 1))) {
      bool infer = false;
      Type ty = parse_type(&infer);
      std::vector<ExprP> args = parse_args();
      return b.construct(ty, infer, args);
    }
    if (name == "modf" && is_punct("(", 1) && !functions.count(name)) {  // modf(e).fract / modf(e).whole
      advance();
      std::vector<ExprP> args = parse_args();
      if (args.size() != 1) b.error("modf takes one argument");
      expect(".");
      const std::string m = expect_ident("fract or whole");
      ExprP x = b.concretize(args[0]);
      ExprP whole = b.call_builtin("trunc", {x});
      if (m == "whole") return whole;
      if (m == "fract") return b.binary(Op::Sub, x, whole);
      b.error("modf() result has members fract and whole");
    }
    if (name == "ldexp" && is_punct("(", 1) && !functions.count(name)) {  // x * 2^e
      advance();
      std::vector<ExprP> args = parse_args();
      if (args.size() != 2) b.error("ldexp takes two arguments");
      ExprP e = b.concretize(args[1]);
      if (!e->ty.is_int()) b.error("second argument of ldexp() must be an integer");
      return b.binary(Op::Mul, args[0], b.call_builtin("exp2", {b.construct(e->ty.with_sk(Sk::F32), false, {e})}));
    }
    if (name == "bitcast" && is_punct("<", 1)) {  // bitcast<T>(e): T = f32 / i32 / u32 or a vector of them
      advance();
      expect("<");
      Type t = parse_type();
      expect_close_angle();
      std::vector<ExprP> args = parse_args();
      if (args.size() != 1) b.error("bitcast takes one argument");
      if (!(t.is_scalar() || t.is_vector())) b.error("bitcast to " + t.str());
      ExprP e = b.bitcast(t.sk, args[0]);
      if (e->ty != t) b.error("bitcast between types of different size: " + args[0]->ty.str() + " -> " + t.str());
      return e;
    }
    if (is_punct("(", 1)) {
      advance();
      std::vector<ExprP> args = parse_args();
      auto it = functions.find(name);
      if (it != functions.end()) {
        if (it->second->is_entry) b.error("entry point '" + name + "' cannot be called");
        return b.call_user(it->second, args);
      }
      ExprP e = b.call_builtin(name, args);
      if (!e) b.error("unknown function '" + name + "'");
      return e;
    }
    b.cur_line = t.line;
    b.error("unknown identifier '" + name + "'");
  }
};

}  // namespace

void parse_wgsl(const std::string& src, const std::vector<std::string>& builtin_fns, Module* out) {
  WgslParser p(out, builtin_fns);
  p.parse(src);
}

}  // namespace s2m_frontend
