// parse_glsl.cpp -- GLSL (fragment shader) front-end: the subset ShaderToy-style SDF code uses.
// Replaces naga's glsl::Frontend behind convert_glsl_to_wgsl (/root/reference/src/shadertoy.rs:169-194).
#include <cstring>

#include "parse.h"
#include "parser_base.h"

namespace s2m_frontend {

namespace {

class GlslParser : public ParserBase {
 public:
  explicit GlslParser(Module* m) : ParserBase(Lang::Glsl, m) {}
  std::map<std::string, std::vector<Function*>> overloads_;

  // Side effects inside expressions (`d += e = map(p)`, `a[i++]`, `while (i++ < n)`): the effect becomes a
  // statement of its own, issued before the statement that contains the expression (GLSL leaves the
  // order of operand evaluation undefined, so any order is a valid one); where GLSL does define an
  // order -- the right side of && and ||, the branches of ?: -- side effects are rejected.
  std::vector<StmtP>* side_ = nullptr;   // where pending effects go; null: not allowed here
  int cond_depth_ = 0;                   // > 0 while parsing a conditionally evaluated operand
  bool leave_postfix_ = false;           // the statement parser itself handles a trailing ++ / --
  int post_tmp_ = 0;
  struct SideScope {
    GlslParser& p;
    std::vector<StmtP> stmts;
    std::vector<StmtP>* saved;
    int saved_depth;
    explicit SideScope(GlslParser& parser) : p(parser), saved(parser.side_), saved_depth(parser.cond_depth_) { p.side_ = &stmts; p.cond_depth_ = 0; }
    ~SideScope() { p.side_ = saved; p.cond_depth_ = saved_depth; }
    void flush_into(const StmtP& blk) { for (const StmtP& st : stmts) blk->body.push_back(st); stmts.clear(); }
  };
  void need_side_context(const char* what) {
    if (!side_) b.unsupported(std::string(what) + " in this position");
    if (cond_depth_ > 0) b.unsupported(std::string(what) + " inside a conditionally evaluated operand (&&, ||, ?:)");
  }

  void parse(const std::string& src) {
    LexOptions lo;
    lo.glsl = true;
    toks = Lexer(src, lo).run();
    push_scope();
    bool have_main = false;
    while (peek().k != Token::End) {
      if (accept(";")) continue;
      parse_external_declaration(&have_main);
    }
    if (!have_main) throw FrontendError(3, "parse error: GLSL shader has no entry point `void main()`");
  }

 private:
  static bool type_from_name(const std::string& s, Type* t) {
    if (s == "void") { *t = Type::void_(); return true; }
    if (s == "float") { *t = Type::scalar(Sk::F32); return true; }
    if (s == "int") { *t = Type::scalar(Sk::I32); return true; }
    if (s == "uint") { *t = Type::scalar(Sk::U32); return true; }
    if (s == "bool") { *t = Type::scalar(Sk::Bool); return true; }
    if (s.size() == 4 && s.compare(0, 3, "vec") == 0 && s[3] >= '2' && s[3] <= '4') { *t = Type::vec(Sk::F32, s[3] - '0'); return true; }
    if (s.size() == 4 && s.compare(0, 3, "mat") == 0 && s[3] >= '2' && s[3] <= '4') { *t = Type::mat(s[3] - '0'); return true; }
    if (s.size() == 6 && s.compare(0, 3, "mat") == 0 && s[4] == 'x' && s[3] == s[5] && s[3] >= '2' && s[3] <= '4') { *t = Type::mat(s[3] - '0'); return true; }
    if (s.size() == 5 && s.compare(1, 3, "vec") == 0 && s[4] >= '2' && s[4] <= '4') {
      const Sk sk = s[0] == 'i' ? Sk::I32 : s[0] == 'u' ? Sk::U32 : s[0] == 'b' ? Sk::Bool : Sk::F32;
      if (s[0] == 'i' || s[0] == 'u' || s[0] == 'b') { *t = Type::vec(sk, s[4] - '0'); return true; }
    }
    return false;
  }
  bool at_type() const {
    if (peek().k != Token::Ident) return false;
    Type t;
    const std::string& s = peek().text;
    if (structs.count(s) && !lookup(s)) return true;
    return type_from_name(s, &t) || s.compare(0, 3, "mat") == 0 || s == "double" || s.compare(0, 4, "dvec") == 0 || s.compare(0, 7, "sampler") == 0;
  }
  // `[N]` / `[]` after a type or a declarator name.  *unsized is set for `[]` (length from the initializer).
  Type parse_array_suffix(Type base, bool* unsized) {
    if (unsized) *unsized = false;
    if (!is_punct("[")) return base;
    advance();
    if (accept("]")) {
      if (!unsized) b.error("array needs a length here");
      *unsized = true;
      if (is_punct("[")) b.unsupported("arrays of arrays");
      return base;
    }
    const int len = array_len(parse_expr());
    expect("]");
    if (is_punct("[")) b.unsupported("arrays of arrays");
    return mod->array_of(base, len);
  }
  Type parse_type(bool* unsized = nullptr) {
    const std::string s = expect_ident("a type");
    Type t;
    if (structs.count(s)) t = Type::struct_(structs[s]);
    else if (!type_from_name(s, &t)) b.unsupported("GLSL type '" + s + "'");
    return parse_array_suffix(t, unsized);
  }
  // declared type of one declarator: the base type plus the declarator's own array suffix;
  // an unsized array takes its length from the initializer
  Type declarator_type(Type base, bool base_unsized, ExprP* init, bool allow_init) {
    bool unsized = base_unsized;
    Type ty = base;
    if (is_punct("[")) {
      if (base.is_array() || base_unsized) b.unsupported("arrays of arrays");
      ty = parse_array_suffix(base, &unsized);
    }
    if (allow_init && accept("=")) {
      if (is_punct("{")) {  // GLSL 4.20 initializer list: float a[3] = {1., 2., 3.};  S s = {1., vec3(0.)};
        *init = parse_initializer_list(unsized ? Type::void_() : ty, base);
      } else {
        *init = parse_assignment_expr();
      }
    }
    if (unsized) {
      if (!*init || !(*init)->ty.is_array() || !((*init)->ty.adef->elem == base)) b.error("unsized array needs an array initializer of the same element type");
      ty = (*init)->ty;
    }
    return ty;
  }
  // `{ a, b, ... }` for an array (elements), a struct (members in order) or a vector / matrix
  // (constructor arguments).  `ty` void: an unsized array of `elem`.
  ExprP parse_initializer_list(Type ty, Type elem) {
    expect("{");
    std::vector<ExprP> items;
    const bool as_array = ty.is_void() || ty.is_array();
    const Type el = ty.is_array() ? ty.adef->elem : elem;
    while (!is_punct("}")) {
      Type item_ty = as_array ? el : ty.is_struct() ? (items.size() < ty.sdef->field_types.size() ? ty.sdef->field_types[items.size()] : Type::void_()) : Type::void_();
      if (is_punct("{")) {
        if (item_ty.is_void() || !(item_ty.is_aggregate() || item_ty.is_vector() || item_ty.is_matrix())) b.error("unexpected nested initializer list");
        items.push_back(parse_initializer_list(item_ty, item_ty.is_array() ? item_ty.adef->elem : item_ty));
      } else {
        items.push_back(parse_assignment_expr());
      }
      if (!accept(",")) break;
    }
    expect("}");
    if (ty.is_void()) {
      if (items.empty()) b.error("cannot infer the length of an empty initializer list");
      ty = mod->array_of(el, (int)items.size());
    }
    return b.construct(ty, false, items);
  }

  void parse_struct() {  // struct Name { float a; vec3 b, c; float d[3]; };
    advance();
    StructDef* d = declare_struct(expect_ident("struct name"));
    expect("{");
    while (!is_punct("}")) {
      parse_qualifiers();
      bool unsized = false;
      Type base = parse_type(&unsized);
      if (unsized) b.error("unsized array member");
      for (;;) {
        const std::string fname = expect_ident("a member name");
        add_field(d, fname, is_punct("[") ? parse_array_suffix(base, nullptr) : base);
        if (!accept(",")) break;
      }
      expect(";");
    }
    expect("}");
    if (d->field_names.empty()) b.error("struct " + d->name + " has no members");
    if (!is_punct(";")) b.unsupported("variable declared together with its struct");
    expect(";");
  }

  struct Qualifiers { bool is_const = false, uniform = false, in = false, out = false, inout = false; };
  Qualifiers parse_qualifiers() {
    Qualifiers q;
    for (;;) {
      if (accept_ident("const")) q.is_const = true;
      else if (accept_ident("uniform")) q.uniform = true;
      else if (accept_ident("in")) q.in = true;
      else if (accept_ident("out")) q.out = true;
      else if (accept_ident("inout")) q.inout = true;
      else if (accept_ident("highp") || accept_ident("mediump") || accept_ident("lowp") || accept_ident("flat") ||
               accept_ident("smooth") || accept_ident("noperspective") || accept_ident("centroid") || accept_ident("invariant") ||
               accept_ident("precise")) {}
      else if (is_ident("layout")) {
        advance();
        expect("(");
        int depth = 1;
        while (depth > 0 && peek().k != Token::End) { if (is_punct("(")) ++depth; if (is_punct(")")) --depth; advance(); }
      } else break;
    }
    return q;
  }

  void parse_external_declaration(bool* have_main) {
    b.cur_line = peek().line;
    if (accept_ident("precision")) { while (!is_punct(";") && peek().k != Token::End) advance(); expect(";"); return; }
    if (is_ident("struct")) { parse_struct(); return; }
    Qualifiers q = parse_qualifiers();
    if (is_ident("buffer") || is_ident("shared")) b.unsupported("storage qualifier " + peek().text);
    if (accept(";")) return;  // e.g. `layout(...) in;`
    if (peek().k == Token::Ident && is_opaque_type(peek().text)) {
      // `uniform sampler2D tex;` -- there is nothing to sample from here: the declaration is accepted, and a
      // function that uses the name is left out like any other that needs a texture
      const std::string tname = advance().text;
      for (;;) {
        const std::string vname = expect_ident("a name");
        while (is_punct("[")) { while (!is_punct("]") && peek().k != Token::End) advance(); expect("]"); }
        opaque_globals_[vname] = tname;
        if (!accept(",")) break;
      }
      expect(";");
      return;
    }
    if (q.uniform && peek().k == Token::Ident && is_punct("{", 1) && !structs.count(peek().text)) { parse_uniform_block(q); return; }
    bool unsized = false;
    Type base = parse_type(&unsized);
    const std::string name = expect_ident("a name");
    if (is_punct("(")) {
      if (unsized) b.error("function returning an unsized array");
      parse_function(base, name, have_main);
      return;
    }
    // global variable(s)
    std::string n = name;
    for (;;) {
      ExprP init;
      const Type ty = declarator_type(base, unsized, &init, true);
      declare_global(q, ty, n, init);
      if (!accept(",")) break;
      n = expect_ident("a name");
    }
    expect(";");
  }

  std::map<std::string, std::string> opaque_globals_;   // name -> type of sampler / image uniforms
  static bool is_opaque_type(const std::string& s) {
    auto starts = [&](const char* p) { return s.compare(0, strlen(p), p) == 0; };
    return starts("sampler") || starts("isampler") || starts("usampler") || starts("image") || starts("iimage") || starts("uimage") ||
           starts("texture") || starts("itexture") || starts("utexture") || s == "atomic_uint" || starts("subpassInput");
  }
  // layout(...) uniform Block { float t; vec3 c; } [instance];  -- like every uniform here the members read as
  // zero.  Without an instance name the members are globals; with one they are the fields of one global struct.
  void parse_uniform_block(const Qualifiers& q) {
    const std::string block = advance().text;
    expect("{");
    std::vector<std::pair<std::string, Type>> members;
    while (!is_punct("}")) {
      parse_qualifiers();
      bool unsized = false;
      Type base = parse_type(&unsized);
      if (unsized) b.error("unsized array member");
      for (;;) {
        const std::string fname = expect_ident("a member name");
        members.emplace_back(fname, is_punct("[") ? parse_array_suffix(base, nullptr) : base);
        if (!accept(",")) break;
      }
      expect(";");
    }
    expect("}");
    if (members.empty()) b.error("uniform block " + block + " has no members");
    if (peek().k == Token::Ident) {
      const std::string inst = advance().text;
      if (is_punct("[")) b.unsupported("arrays of uniform blocks");
      StructDef* d = declare_struct(block);
      for (const auto& m : members) add_field(d, m.first, m.second);
      declare_global(q, Type::struct_(d), inst, nullptr);
    } else {
      for (const auto& m : members) declare_global(q, m.second, m.first, nullptr);
    }
    expect(";");
  }

  void declare_global(const Qualifiers& q, Type ty, const std::string& name, ExprP init) {
    if (ty.is_void()) b.error("variable of type void");
    const bool resource = q.uniform || q.in || q.out;
    if (resource) init = nullptr;
    if (init) init = b.coerce(init, ty, "initializer");
    Var* v = declare(name, ty, q.is_const ? Var::ModuleConst : Var::Global);
    v->immutable = q.is_const || q.uniform || q.in;
    if (init) {
      ConstVal cv;
      if (b.const_eval(*init, &cv)) { v->has_const = true; v->cval = cv; }
      mod->global_init[v] = init;
    } else {
      if (q.is_const) b.error("const '" + name + "' needs an initializer");
      v->has_const = true;
      v->cval.ty = ty;
    }
    mod->globals.push_back(v);
  }

  void parse_function(Type ret, const std::string& name, bool* have_main) {
    std::vector<Var*> params;
    expect("(");
    if (is_ident("void") && is_punct(")", 1)) advance();
    while (!is_punct(")")) {
      Qualifiers q = parse_qualifiers();
      Type t = parse_type();
      if (t.is_void()) b.error("parameter of type void");
      std::string pn = "_p" + std::to_string(params.size());
      if (peek().k == Token::Ident) pn = advance().text;
      if (is_punct("[")) { if (t.is_array()) b.unsupported("arrays of arrays"); t = parse_array_suffix(t, nullptr); }
      Var* v = mod->new_var();
      v->name = pn; v->ty = t; v->storage = Var::Param;
      v->immutable = q.is_const;
      v->by_ref = q.out || q.inout;
      params.push_back(v);
      if (!accept(",")) break;
    }
    expect(")");
    // GLSL allows overloading: functions with the same name are distinguished by parameter types.
    // The first keeps its name, later ones are emitted as name_1, name_2, ... (like naga).
    Function* fn = nullptr;
    std::vector<Function*>& set = overloads_[name];
    for (Function* f : set) {
      bool same = f->params.size() == params.size();
      for (size_t i = 0; same && i < params.size(); ++i) same = f->params[i]->ty == params[i]->ty && f->params[i]->by_ref == params[i]->by_ref;
      if (same) { fn = f; break; }
    }
    if (fn) {
      if (fn->ret != ret) b.error("function '" + name + "' redeclared with a different return type");
      if (fn->body && is_punct("{")) b.error("redefinition of function '" + name + "'");
    } else {
      if (scopes[0].count(name)) b.error("'" + name + "' redeclared as a function");
      mod->functions.emplace_back(new Function());
      fn = mod->functions.back().get();
      fn->name = name;
      for (int k = 1; functions.count(fn->name) || scopes[0].count(fn->name); ++k) fn->name = name + "_" + std::to_string(k);
      fn->ret = ret; fn->line = peek().line;
      functions[fn->name] = fn;
      set.push_back(fn);
    }
    if (accept(";")) { if (fn->params.empty()) fn->params = params; return; }  // prototype
    fn->params = params;
    if (name == "main") {
      *have_main = true;
      fn->is_entry = true;
      skip_braces();  // the fragment entry point is discarded by the caller (shader.rs:84-86)
      return;
    }
    if (name == "mainImage") {  // ShaderToy image entry: never part of the SDF; keep the signature only
      fn->is_entry = false;
      skip_braces();
      fn->body = mk_stmt(Stmt::Block);
      return;
    }
    // A ShaderToy image pass is mostly shading code around the distance function.  The reference hands
    // all of it to naga (shadertoy.rs:154-167) and only ever calls the SDF; here a function whose body
    // needs something with no counterpart in this engine (textures, derivatives, ...) is left out, and
    // the error is raised only if something that is kept calls it.
    const size_t body_tok = pos;
    const size_t scope_depth = scopes.size();
    cur_fn = fn;
    try {
      push_scope();
      for (Var* p : params) {
        if (scopes.back().count(p->name)) b.error("duplicate parameter '" + p->name + "'");
        scopes.back()[p->name] = p;
      }
      fn->body = parse_compound(false);
      pop_scope();
    } catch (const FrontendError& e) {
      if (e.status != 11) throw;
      while (scopes.size() > scope_depth) pop_scope();
      side_ = nullptr; cond_depth_ = 0; leave_postfix_ = false;
      pos = body_tok;
      skip_braces();
      fn->body = nullptr;
      dropped_[fn] = e.what();
      mod->dropped.emplace_back(fn->name, e.what());
    }
    cur_fn = nullptr;
  }
  std::map<const Function*, std::string> dropped_;

  // ---------------------------------------------------------------- statements
  StmtP parse_compound(bool new_scope) {
    expect("{");
    StmtP blk = mk_stmt(Stmt::Block);
    if (new_scope) push_scope();
    while (!is_punct("}")) {
      if (peek().k == Token::End) perr("unterminated block");
      parse_statement_into(blk);
    }
    expect("}");
    if (new_scope) pop_scope();
    return blk;
  }
  // a statement used as a loop / if body: always returns a Block
  StmtP parse_body() {
    if (is_punct("{")) return parse_compound(true);
    StmtP blk = mk_stmt(Stmt::Block);
    push_scope();
    parse_statement_into(blk);
    pop_scope();
    return blk;
  }

  void parse_declaration_into(StmtP blk) {
    Qualifiers q = parse_qualifiers();
    bool unsized = false;
    Type base = parse_type(&unsized);
    if (base.is_void()) b.error("variable of type void");
    for (;;) {
      const std::string name = expect_ident("a variable name");
      ExprP init;
      SideScope sc(*this);
      const Type ty = declarator_type(base, unsized, &init, true);
      sc.flush_into(blk);
      if (init) init = b.coerce(init, ty, "initializer");
      Var* v = declare(name, ty, Var::Local);
      v->immutable = q.is_const;
      if (q.is_const && !init) b.error("const '" + name + "' needs an initializer");
      StmtP s = mk_stmt(Stmt::VarDecl);
      s->var = v; s->a = init;
      blk->body.push_back(s);
      if (!accept(",")) break;
    }
  }

  // expression statement without ';' : assignment, ++/--, call
  StmtP parse_expression_statement() {
    if (is_punct("++") || is_punct("--")) {
      const bool inc = advance().text == "++";
      ExprP lhs = parse_unary();
      return make_assign(lhs, b.binary(inc ? Op::Add : Op::Sub, lhs, one_for(lhs)));
    }
    leave_postfix_ = true;
    ExprP lhs = parse_unary();
    leave_postfix_ = false;
    if (is_punct("=")) {
      advance();
      return make_assign(lhs, parse_assignment_expr());
    }
    if (peek().k == Token::Punct) {
      const std::string p = peek().text;
      if (p == "+=" || p == "-=" || p == "*=" || p == "/=" || p == "%=" || p == "&=" || p == "|=" || p == "^=" || p == "<<=" || p == ">>=") {
        advance();
        ExprP rhs = parse_assignment_expr();
        return make_assign(lhs, b.binary(compound_op(p), lhs, rhs));
      }
      if (p == "++" || p == "--") {
        advance();
        return make_assign(lhs, b.binary(p == "++" ? Op::Add : Op::Sub, lhs, one_for(lhs)));
      }
    }
    if (lhs->k == Expr::UserCall || lhs->k == Expr::Call) {
      StmtP s = mk_stmt(Stmt::CallStmt);
      s->a = lhs;
      return s;
    }
    perr("expected a statement (expression statements without effect are not supported)");
  }
  ExprP one_for(const ExprP& lhs) {
    if (lhs->ty.is_float()) return b.lit_float(1.0, Sk::F32);
    return b.lit_int(1, lhs->ty.sk);
  }

  ExprP parse_condition() {
    ExprP c = parse_expr();
    if (!c->ty.is_bool() || !c->ty.is_scalar()) b.error("condition must be a bool, found " + c->ty.str());
    return c;
  }

  void parse_statement_into(StmtP blk) {
    b.cur_line = peek().line;
    if (accept(";")) return;
    if (is_punct("{")) { blk->body.push_back(parse_compound(true)); return; }
    if (is_ident("const") || is_ident("highp") || is_ident("mediump") || is_ident("lowp") || (at_type() && (peek(1).k == Token::Ident || is_punct("[", 1)))) {
      parse_declaration_into(blk);
      expect(";");
      return;
    }
    if (accept_ident("return")) {
      StmtP s = mk_stmt(Stmt::Return);
      if (!is_punct(";")) {
        if (cur_fn->ret.is_void()) b.error("return with a value in a void function");
        SideScope sc(*this);
        s->a = b.coerce(parse_expr(), cur_fn->ret, "return");
        sc.flush_into(blk);
      } else if (!cur_fn->ret.is_void()) b.error("return without a value");
      expect(";");
      blk->body.push_back(s);
      return;
    }
    if (accept_ident("if")) { blk->body.push_back(parse_if(blk)); return; }
    if (accept_ident("for")) {
      StmtP s = mk_stmt(Stmt::For);
      push_scope();
      expect("(");
      std::vector<StmtP> inits;  // several declarators / comma-separated init expressions (+ their side effects)
      if (!is_punct(";")) {
        if (at_type() && (peek(1).k == Token::Ident || is_punct("[", 1))) {
          StmtP tmp = mk_stmt(Stmt::Block);
          parse_declaration_into(tmp);
          inits = tmp->body;
        } else {
          for (;;) {
            SideScope sc(*this);
            StmtP st = parse_expression_statement();
            for (const StmtP& e : sc.stmts) inits.push_back(e);
            inits.push_back(st);
            if (!accept(",")) break;
          }
        }
      }
      expect(";");
      std::vector<StmtP> cond_effects;
      if (!is_punct(";")) {
        SideScope sc(*this);
        s->a = parse_condition();
        cond_effects = sc.stmts;
      }
      expect(";");
      std::vector<StmtP> conts;
      if (!is_punct(")")) {
        for (;;) {
          SideScope sc(*this);
          StmtP st = parse_expression_statement();
          for (const StmtP& e : sc.stmts) conts.push_back(e);
          conts.push_back(st);
          if (!accept(",")) break;
        }
      }
      expect(")");
      ++loop_depth;
      StmtP body = parse_body();
      --loop_depth;
      pop_scope();
      if (inits.size() <= 1 && conts.size() <= 1 && cond_effects.empty()) {
        if (!inits.empty()) s->init = inits[0];
        if (!conts.empty()) s->cont = conts[0];
        s->body.push_back(body);
        blk->body.push_back(s);
        return;
      }
      // for (a, b; c; d, e) body  ==  { a; b; loop { [effects of c;] if (!c) break; body; continuing { d; e; } } }
      // (`continue` in the body still runs d and e: that is what a WGSL continuing block does)
      StmtP outer = mk_stmt(Stmt::Block);
      for (const StmtP& i : inits) outer->body.push_back(i);
      outer->body.push_back(make_loop(cond_effects, s->a, body, conts));
      blk->body.push_back(outer);
      return;
    }
    if (accept_ident("while")) {
      StmtP s = mk_stmt(Stmt::While);
      expect("(");
      std::vector<StmtP> cond_effects;
      {
        SideScope sc(*this);
        s->a = parse_condition();
        cond_effects = sc.stmts;
      }
      expect(")");
      ++loop_depth;
      StmtP body = parse_body();
      --loop_depth;
      if (cond_effects.empty()) s->body.push_back(body);
      else s = make_loop(cond_effects, s->a, body, {});   // while (i++ < n): the effect runs before every test
      blk->body.push_back(s);
      return;
    }
    if (accept_ident("do")) {
      StmtP s = mk_stmt(Stmt::DoWhile);
      ++loop_depth;
      StmtP body = parse_body();
      --loop_depth;
      if (!accept_ident("while")) perr("expected 'while' after do body");
      expect("(");
      std::vector<StmtP> cond_effects;
      {
        SideScope sc(*this);
        s->a = parse_condition();
        cond_effects = sc.stmts;
      }
      expect(")");
      expect(";");
      if (cond_effects.empty()) {
        s->body.push_back(body);
      } else {  // loop { body; continuing { effects; break if !cond; } }
        StmtP loop = mk_stmt(Stmt::Loop);
        loop->body.push_back(body);
        loop->cont = mk_stmt(Stmt::Block);
        for (const StmtP& e : cond_effects) loop->cont->body.push_back(e);
        loop->break_if = b.unary(Op::Not, s->a);
        s = loop;
      }
      blk->body.push_back(s);
      return;
    }
    if (accept_ident("break")) { if (!loop_depth && !switch_depth) b.error("break outside of a loop or switch"); expect(";"); blk->body.push_back(mk_stmt(Stmt::Break)); return; }
    if (accept_ident("continue")) { if (!loop_depth) b.error("continue outside of a loop"); expect(";"); blk->body.push_back(mk_stmt(Stmt::Continue)); return; }
    if (accept_ident("discard")) { expect(";"); blk->body.push_back(mk_stmt(Stmt::Discard)); return; }
    if (is_ident("precision") && peek(1).k == Token::Ident) { while (!is_punct(";") && peek().k != Token::End) advance(); expect(";"); return; }  // precision highp float;
    if (accept_ident("switch")) { blk->body.push_back(parse_switch(blk)); return; }
    SideScope sc(*this);
    for (;;) {  // `a = 1., b = 2.;` -- the comma operator at statement level is a sequence of statements
      StmtP s = parse_expression_statement();
      sc.flush_into(blk);
      blk->body.push_back(s);
      if (!accept(",")) break;
      b.cur_line = peek().line;
    }
    expect(";");
  }

  static bool ends_flow(const Stmt& blk) {  // last statement leaves the case for good
    if (blk.body.empty()) return false;
    const Stmt& last = *blk.body.back();
    const Stmt::K k = last.k;
    if (k == Stmt::Block) return ends_flow(last);   // case 1: { ...; return x; }
    if (k == Stmt::If) return last.then_s && last.else_s && ends_flow_stmt(*last.then_s) && ends_flow_stmt(*last.else_s);
    return k == Stmt::Break || k == Stmt::Return || k == Stmt::Continue || k == Stmt::Discard;
  }
  static bool ends_flow_stmt(const Stmt& s) {
    if (s.k == Stmt::Block) return ends_flow(s);
    if (s.k == Stmt::If) return s.then_s && s.else_s && ends_flow_stmt(*s.then_s) && ends_flow_stmt(*s.else_s);
    return s.k == Stmt::Break || s.k == Stmt::Return || s.k == Stmt::Continue || s.k == Stmt::Discard;
  }
  // switch (e) { case 1: case 2: ...; break; default: ... }  -- labels group into cases; a case
  // that runs into the next label without break/return/continue/discard would fall through,
  // which the IR (like WGSL) does not model.
  StmtP parse_switch(const StmtP& blk) {
    StmtP sw = mk_stmt(Stmt::Switch);
    expect("(");
    {
      SideScope sc(*this);
      sw->a = switch_selector(parse_expr());
      sc.flush_into(blk);
    }
    expect(")");
    expect("{");
    ++switch_depth;
    push_scope();
    // Groups of labels and the statements that follow them.  The IR (like WGSL) has no fall-through: a group that
    // runs into the next label without break / return / continue / discard gets the statements of the following
    // group(s) parsed into it again, up to the first one that leaves -- which is what falling through executes.
    struct Group { size_t labels = 0, stmts = 0, end = 0; };
    std::vector<Group> groups;
    {
      int depth = 0;
      size_t i = pos;
      bool in_labels = false;
      for (;; ++i) {
        const Token& t = toks[i];
        if (t.k == Token::End) perr("unterminated switch");
        if (t.k == Token::Punct && (t.text == "{" || t.text == "(" || t.text == "[")) ++depth;
        if (t.k == Token::Punct && (t.text == "}" || t.text == ")" || t.text == "]")) { if (depth == 0) break; --depth; }
        if (depth == 0 && t.k == Token::Ident && (t.text == "case" || t.text == "default")) {
          if (!in_labels) { if (!groups.empty()) groups.back().end = i; groups.push_back(Group()); groups.back().labels = i; in_labels = true; }
          int ternaries = 0;   // skip to the label's own ':'
          for (++i;; ++i) {
            const Token& u = toks[i];
            if (u.k == Token::End) perr("unterminated case label");
            if (u.k == Token::Punct && u.text == "?") ++ternaries;
            if (u.k == Token::Punct && u.text == ":") { if (ternaries == 0) break; --ternaries; }
          }
          groups.back().stmts = i + 1;
          continue;
        }
        if (groups.empty()) perr("statement before the first case label");
        in_labels = false;
      }
      if (!groups.empty()) groups.back().end = i;
      for (Group& g : groups) if (g.end < g.stmts) g.end = g.stmts;
      // parse group by group; `i` is the closing brace
      for (size_t g = 0; g < groups.size(); ++g) {
        StmtP cur = mk_stmt(Stmt::Case);
        cur->body.push_back(mk_stmt(Stmt::Block));
        sw->body.push_back(cur);
        pos = groups[g].labels;
        while (pos < groups[g].stmts) {
          if (accept_ident("default")) cur->is_default = true;
          else { if (!accept_ident("case")) perr("expected a case label"); cur->case_values.push_back(case_value(parse_binary(0), sw->a->ty)); }
          expect(":");
        }
        while (pos < groups[g].end) parse_statement_into(cur->body[0]);
        for (size_t h = g + 1; h < groups.size() && !ends_flow(*cur->body[0]); ++h) {
          push_scope();   // the statements are parsed a second time: their declarations are new variables
          pos = groups[h].stmts;
          while (pos < groups[h].end) parse_statement_into(cur->body[0]);
          pop_scope();
        }
      }
      pos = i;
    }
    expect("}");
    pop_scope();
    --switch_depth;
    for (const StmtP& c : sw->body) {  // the closing break of a case is implied by the IR
      std::vector<StmtP>& bd = c->body[0]->body;
      if (!bd.empty() && bd.back()->k == Stmt::Break) bd.pop_back();
    }
    check_cases(*sw);
    return sw;
  }

  // loop { effects; if (!cond) break; body; continuing { conts } }
  StmtP make_loop(const std::vector<StmtP>& effects, const ExprP& cond, const StmtP& body, const std::vector<StmtP>& conts) {
    StmtP loop = mk_stmt(Stmt::Loop);
    StmtP lbody = mk_stmt(Stmt::Block);
    for (const StmtP& e : effects) lbody->body.push_back(e);
    if (cond) {
      StmtP guard = mk_stmt(Stmt::If);
      guard->a = b.unary(Op::Not, cond);
      guard->then_s = mk_stmt(Stmt::Block);
      guard->then_s->body.push_back(mk_stmt(Stmt::Break));
      lbody->body.push_back(guard);
    }
    lbody->body.push_back(body);
    loop->body.push_back(lbody);
    if (!conts.empty()) {
      loop->cont = mk_stmt(Stmt::Block);
      for (const StmtP& c : conts) loop->cont->body.push_back(c);
    }
    return loop;
  }

  // `blk`: where side effects of the condition go (before the if); null for `else if` chains, whose
  // conditions are evaluated conditionally
  StmtP parse_if(const StmtP& blk) {
    StmtP s = mk_stmt(Stmt::If);
    expect("(");
    if (blk) {
      SideScope sc(*this);
      s->a = parse_condition();
      sc.flush_into(blk);
    } else {
      std::vector<StmtP>* saved = side_;
      side_ = nullptr;
      s->a = parse_condition();
      side_ = saved;
    }
    expect(")");
    s->then_s = parse_body();
    if (accept_ident("else")) {
      if (accept_ident("if")) s->else_s = parse_if(nullptr);
      else s->else_s = parse_body();
    }
    return s;
  }

  // ---------------------------------------------------------------- expressions
  static bool has_user_call(const Expr& e) {
    if (e.k == Expr::UserCall) return true;
    for (const ExprP& a : e.args) if (a && has_user_call(*a)) return true;
    return false;
  }
  // c ? t : f -- GLSL evaluates exactly one of the arms.  An arm that calls a user function (it may
  // assign module-scope variables or write through an out parameter, and it may be expensive) or that
  // contains an assignment / ++ / -- is therefore lowered the way naga lowers it: a temporary, declared
  // in front of the statement, assigned in an if / else whose branches hold that arm's effects.  Arms
  // made of operators and builtins only stay an expression (WGSL select(): both sides evaluated, which
  // is unobservable for pure arms).  Where no statement can be issued (conditions of `else if`, the
  // right side of && and ||, an enclosing ?: that stayed an expression) calls stay eager as well;
  // check_eager_conditionals() rejects the module afterwards if such a call has a side effect.
  ExprP parse_conditional_arms(ExprP c) {
    const bool can_lower = side_ != nullptr && cond_depth_ == 0;
    if (!can_lower) {
      ++cond_depth_;
      ExprP t = parse_assignment_expr();
      expect(":");
      ExprP f = parse_assignment_expr();
      --cond_depth_;
      return b.ternary(c, t, f);
    }
    std::vector<StmtP> t_side, f_side;
    ExprP t, f;
    {
      SideScope sc(*this);
      t = parse_assignment_expr();
      t_side.swap(sc.stmts);
    }
    expect(":");
    {
      SideScope sc(*this);
      f = parse_assignment_expr();
      f_side.swap(sc.stmts);
    }
    ExprP sel = b.ternary(c, t, f);   // type checks and coerces the arms to their common type
    if (sel->k != Expr::Ternary) return sel;   // folded (constant condition)
    if (t_side.empty() && f_side.empty() && !has_user_call(*sel->args[1]) && !has_user_call(*sel->args[2])) return sel;
    Var* tmp = declare("_cond" + std::to_string(++post_tmp_), sel->ty, Var::Local);
    StmtP decl = mk_stmt(Stmt::VarDecl);
    decl->var = tmp;
    side_->push_back(decl);
    StmtP br = mk_stmt(Stmt::If);
    br->a = sel->args[0];
    br->then_s = mk_stmt(Stmt::Block);
    for (const StmtP& st : t_side) br->then_s->body.push_back(st);
    br->then_s->body.push_back(make_assign(b.var_ref(tmp), sel->args[1]));
    br->else_s = mk_stmt(Stmt::Block);
    for (const StmtP& st : f_side) br->else_s->body.push_back(st);
    br->else_s->body.push_back(make_assign(b.var_ref(tmp), sel->args[2]));
    side_->push_back(br);
    return b.var_ref(tmp);
  }
  ExprP parse_expr() { return parse_assignment_expr(); }
  ExprP parse_assignment_expr() {  // the ?: level; an assignment here is a side effect of the enclosing statement
    ExprP c = parse_binary(0);
    if (accept("?")) return parse_conditional_arms(c);
    if (peek().k == Token::Punct) {
      const std::string p = peek().text;
      const bool compound = p == "+=" || p == "-=" || p == "*=" || p == "/=" || p == "%=" || p == "&=" || p == "|=" || p == "^=" || p == "<<=" || p == ">>=";
      if (p == "=" || compound) {
        need_side_context("an assignment inside an expression");
        advance();
        ExprP rhs = parse_assignment_expr();   // right-associative: a = b = c
        side_->push_back(make_assign(c, compound ? b.binary(compound_op(p), c, rhs) : rhs));
        return c;                               // the value of the assignment is the assigned variable
      }
    }
    return c;
  }
  static int prec_of(const std::string& p) {
    if (p == "||") return 1;
    if (p == "^^") return 2;
    if (p == "&&") return 3;
    if (p == "|") return 4;
    if (p == "^") return 5;
    if (p == "&") return 6;
    if (p == "==" || p == "!=") return 7;
    if (p == "<" || p == ">" || p == "<=" || p == ">=") return 8;
    if (p == "<<" || p == ">>") return 9;
    if (p == "+" || p == "-") return 10;
    if (p == "*" || p == "/" || p == "%") return 11;
    return -1;
  }
  ExprP parse_binary(int min_prec) {
    ExprP lhs = parse_unary();
    for (;;) {
      if (peek().k != Token::Punct) break;
      std::string p = peek().text;
      const bool logical_xor = p == "^" && is_punct("^", 1);  // the lexer has no ^^ token
      if (logical_xor) p = "^^";
      const int prec = prec_of(p);
      if (prec < 0 || prec < min_prec) break;
      advance();
      if (logical_xor) advance();
      const bool short_circuit = p == "&&" || p == "||";
      if (short_circuit) ++cond_depth_;
      ExprP rhs = parse_binary(prec + 1);
      if (short_circuit) --cond_depth_;
      if (logical_xor) {
        if (!lhs->ty.is_bool() || !rhs->ty.is_bool() || !lhs->ty.is_scalar() || !rhs->ty.is_scalar()) b.error("^^ needs bool operands");
        lhs = b.binary(Op::Ne, lhs, rhs);
        continue;
      }
      Op op = p == "||" ? Op::Or : p == "&&" ? Op::And : p == "|" ? Op::BitOr : p == "^" ? Op::BitXor : p == "&" ? Op::BitAnd
              : p == "==" ? Op::Eq : p == "!=" ? Op::Ne : p == "<" ? Op::Lt : p == ">" ? Op::Gt : p == "<=" ? Op::Le : p == ">=" ? Op::Ge
              : p == "<<" ? Op::Shl : p == ">>" ? Op::Shr : p == "+" ? Op::Add : p == "-" ? Op::Sub : p == "*" ? Op::Mul : p == "/" ? Op::Div : Op::Rem;
      if (op == Op::Rem && (lhs->ty.is_float() || rhs->ty.is_float())) b.error("% needs integer operands in GLSL (use mod())");
      if ((op == Op::Eq || op == Op::Ne) && (lhs->ty.is_matrix() || rhs->ty.is_matrix())) {
        // GLSL: == and != compare whole operands and yield ONE bool (equal() / notEqual() are the component-wise forms)
        if (lhs->ty != rhs->ty) b.error("== / != on matrices of different sizes");
        ExprP acc;
        for (int c = 0; c < lhs->ty.n; ++c) {
          ExprP col = b.binary(op, b.index(lhs, b.lit_int(c, Sk::I32)), b.index(rhs, b.lit_int(c, Sk::I32)));
          col = b.call_builtin(op == Op::Eq ? "all" : "any", {col});
          acc = acc ? b.binary(op == Op::Eq ? Op::And : Op::Or, acc, col) : col;
        }
        lhs = acc;
        continue;
      }
      lhs = b.binary(op, lhs, rhs);
      if ((op == Op::Eq || op == Op::Ne) && lhs->ty.is_vector()) lhs = b.call_builtin(op == Op::Eq ? "all" : "any", {lhs});
    }
    return lhs;
  }
  ExprP parse_unary() {
    b.cur_line = peek().line;
    const bool leave = leave_postfix_;   // set by parse_expression_statement for its own first operand only
    leave_postfix_ = false;
    if (accept("-")) return b.unary(Op::Neg, parse_unary());
    if (accept("+")) return parse_unary();
    if (accept("!")) return b.unary(Op::Not, parse_unary());
    if (accept("~")) return b.unary(Op::BitNot, parse_unary());
    if (is_punct("++") || is_punct("--")) {   // ++x inside an expression: x = x + 1 first, value x
      need_side_context("++/-- inside an expression");
      const bool inc = advance().text == "++";
      ExprP lv = parse_unary();
      side_->push_back(make_assign(lv, b.binary(inc ? Op::Add : Op::Sub, lv, one_for(lv))));
      return lv;
    }
    ExprP e = parse_postfix(parse_primary());
    if (is_punct("++") || is_punct("--")) {
      if (leave && (is_punct(";", 1) || is_punct(")", 1) || is_punct(",", 1))) return e;   // `x++;` is the statement itself
      // x++ inside an expression: t = x; x = x + 1; value t
      need_side_context("++/-- inside an expression");
      const bool inc = advance().text == "++";
      if (!Builder::is_lvalue(*e)) b.error("++/-- needs a variable");
      Var* t = declare("_post" + std::to_string(++post_tmp_), e->ty, Var::Local);
      StmtP decl = mk_stmt(Stmt::VarDecl);
      decl->var = t; decl->a = e;
      side_->push_back(decl);
      side_->push_back(make_assign(e, b.binary(inc ? Op::Add : Op::Sub, e, one_for(e))));
      return b.var_ref(t);
    }
    return e;
  }
  ExprP parse_postfix(ExprP e) {
    for (;;) {
      if (accept(".")) {
        const std::string m = expect_ident("a member name");
        if (is_punct("(")) {
          if (m != "length" || !e->ty.is_array()) b.unsupported("method call ." + m + "()");
          advance();
          expect(")");
          e = b.array_length(e);
          continue;
        }
        e = b.member(e, m);
        continue;
      }
      if (is_punct("[")) {
        advance();
        ExprP idx = parse_expr();
        expect("]");
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
    if (is_ident("void") && is_punct(")", 1)) advance();
    while (!is_punct(")")) {
      args.push_back(parse_assignment_expr());
      if (!accept(",")) break;
    }
    expect(")");
    return args;
  }
  ExprP parse_primary() {
    const Token& t = peek();
    b.cur_line = t.line;
    if (t.k == Token::Float) { advance(); return b.lit_float(t.fval, Sk::F32); }
    if (t.k == Token::Int) {
      advance();
      if (t.suffix == 'f') return b.lit_float((double)t.ival, Sk::F32);
      // GLSL 4.1.3: a literal whose bit pattern does not fit in 32 bits is an error; the pattern is used unmodified, so
      // an unsuffixed literal with the sign bit set is a negative int (0xFFFFFFFF == -1)
      if (t.ival < 0 || t.ival > 0xFFFFFFFFll) perr("integer literal '" + t.text + "' does not fit in 32 bits");
      if (t.suffix == 'u') return b.lit_int(t.ival, Sk::U32);
      return b.lit_int(t.ival > 0x7FFFFFFFll ? t.ival - 0x100000000ll : t.ival, Sk::I32);
    }
    if (accept("(")) {
      ExprP e = parse_expr();
      expect(")");
      return e;
    }
    if (t.k != Token::Ident) perr("expected an expression");
    const std::string name = t.text;
    if (name == "true" || name == "false") { advance(); return b.lit_bool(name == "true"); }
    Type ty;
    const bool type_like = (structs.count(name) && !lookup(name)) || type_from_name(name, &ty);
    if (type_like && is_punct("[", 1)) {  // array constructor: float[3](a, b, c) / vec2[](a, b)
      bool unsized = false;
      Type at = parse_type(&unsized);
      if (!is_punct("(")) perr("expected '(' after an array type");
      std::vector<ExprP> args = parse_args();
      if (unsized) {
        if (args.empty()) b.error("cannot infer the length of an empty array constructor");
        at = mod->array_of(at, (int)args.size());
      }
      return b.construct(at, false, args);
    }
    if (structs.count(name) && !lookup(name) && is_punct("(", 1)) {
      advance();
      std::vector<ExprP> args = parse_args();
      return b.construct(Type::struct_(structs[name]), false, args);
    }
    if (is_punct("(", 1) && (type_from_name(name, &ty) || name.compare(0, 3, "mat") == 0)) {
      if (!type_from_name(name, &ty)) b.unsupported("matrix type " + name + " (only square matrices)");
      advance();
      std::vector<ExprP> args = parse_args();
      return b.construct(ty, false, args);
    }
    if (is_punct("(", 1)) {
      // a local variable cannot be called; functions and builtins share one namespace
      advance();
      std::vector<ExprP> args = parse_args();
      auto ov = overloads_.find(name);
      if (ov != overloads_.end() && !ov->second.empty()) {
        Function* best = nullptr;
        int best_score = -1;
        for (Function* f : ov->second) {
          if (f->params.size() != args.size()) continue;
          int score = 0;
          bool ok = true;
          for (size_t i = 0; ok && i < args.size(); ++i) {
            const Type& pt = f->params[i]->ty;
            const Type& at = args[i]->ty;
            if (pt == at) score += 2;
            else if (!f->params[i]->by_ref && pt.k == at.k && pt.n == at.n && pt.sk == Sk::F32 && at.is_int()) score += 1;  // int -> float
            else ok = false;
          }
          if (ok && score > best_score) { best = f; best_score = score; }
        }
        if (!best) {
          if (ov->second.size() == 1) best = ov->second[0];  // let call_user report the precise mismatch
          else b.error("no overload of '" + name + "' matches the argument types");
        }
        if (best->is_entry) b.error("main() cannot be called");
        if (dropped_.count(best)) throw FrontendError(11, "calls '" + best->name + "', which was left out: " + dropped_[best]);
        return b.call_user(best, args);
      }
      if (name == "texture" || name == "texelFetch" || name == "textureLod" || name == "textureGrad" || name == "textureSize" || name == "texture2D")
        b.unsupported("texture sampling (" + name + ")");
      if (name == "dFdx" || name == "dFdy" || name == "fwidth") b.unsupported("screen-space derivatives (" + name + ") outside a fragment stage");
      // bit reinterpretation, vector relational functions, mix with a bool selector
      if (args.size() == 1 && (name == "floatBitsToInt" || name == "floatBitsToUint" || name == "intBitsToFloat" || name == "uintBitsToFloat")) {
        const bool from_float = name[0] == 'f';
        if (from_float != args[0]->ty.is_float()) b.error(name + "() argument has the wrong type: " + args[0]->ty.str());
        return b.bitcast(name == "floatBitsToInt" ? Sk::I32 : name == "floatBitsToUint" ? Sk::U32 : Sk::F32, args[0]);
      }
      if (args.size() == 2) {
        static const std::map<std::string, Op> rel = {{"lessThan", Op::Lt}, {"lessThanEqual", Op::Le}, {"greaterThan", Op::Gt},
                                                       {"greaterThanEqual", Op::Ge}, {"equal", Op::Eq}, {"notEqual", Op::Ne}};
        auto it = rel.find(name);
        if (it != rel.end()) {
          if (!args[0]->ty.is_vector() || !args[1]->ty.is_vector()) b.error(name + "() needs vector operands");
          return b.binary(it->second, args[0], args[1]);
        }
      }
      if (args.size() == 2 && name == "modf") {  // modf(x, out whole): whole = trunc(x) (an effect of the statement), value x - trunc(x)
        need_side_context("modf() with its out parameter");
        if (!Builder::is_lvalue(*args[1])) b.error("second argument of modf() must be a variable");
        ExprP whole = b.call_builtin("trunc", {args[0]});
        side_->push_back(make_assign(args[1], whole));
        return b.binary(Op::Sub, args[0], b.call_builtin("trunc", {args[0]}));
      }
      if (args.size() == 2 && name == "ldexp") {  // x * 2^e
        ExprP e = args[1];
        if (!e->ty.is_int()) b.error("second argument of ldexp() must be an integer");
        return b.binary(Op::Mul, args[0], b.call_builtin("exp2", {b.construct(e->ty.with_sk(Sk::F32), false, {e})}));
      }
      if (args.size() == 1 && (name == "isnan" || name == "isinf")) {  // component-wise; WGSL has neither, so they are spelled out
        ExprP x = args[0];
        if (!x->ty.is_float() || x->ty.is_matrix()) b.error(name + "() needs a float scalar or vector");
        if (name == "isnan") return b.binary(Op::Ne, x, x);
        ExprP big = b.lit_float(3.4028234663852886e38, Sk::F32);   // the largest finite f32
        if (x->ty.is_vector()) big = b.construct(x->ty, false, {big});
        return b.binary(Op::Gt, b.call_builtin("abs", {x}), big);
      }
      if (args.size() == 1 && name == "not") {
        if (!args[0]->ty.is_vector() || !args[0]->ty.is_bool()) b.error("not() needs a bool vector");
        return b.unary(Op::Not, args[0]);
      }
      if (args.size() == 3 && name == "mix" && args[2]->ty.is_bool()) {  // mix(x, y, bool / bvec): component selection
        ExprP e = b.call_builtin("select", {args[0], args[1], args[2]});
        if (!e) b.error("mix() with a bool selector: bad operands");
        return e;
      }
      ExprP e = b.call_builtin(name, args);
      if (!e) b.error("unknown function '" + name + "'");
      if ((name == "bitCount" || name == "findMSB" || name == "findLSB") && e->ty.sk == Sk::U32)
        e = b.bitcast(Sk::I32, e);   // genIType results in GLSL, whatever the argument; same bits (findMSB(0u) = -1)
      return e;
    }
    if (Var* v = lookup(name)) { advance(); return b.var_ref(v); }
    if (name.compare(0, 3, "gl_") == 0) b.unsupported("built-in variable " + name);
    if (name.compare(0, 8, "iChannel") == 0) b.unsupported("ShaderToy channel input " + name);
    if (opaque_globals_.count(name)) b.unsupported(opaque_globals_[name] + " uniform '" + name + "' (no textures or images here)");
    b.error("unknown identifier '" + name + "'");
  }
};

}  // namespace

void parse_glsl(const std::string& src, Module* out) {
  GlslParser p(out);
  p.parse(src);
}

}  // namespace s2m_frontend
