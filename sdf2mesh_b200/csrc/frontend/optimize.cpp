// optimize.cpp -- semantics-preserving IR rewrites applied before CUDA emission.
//
// continue_to_break: in
//     for (int i = 0; i < N; ++i) { P; if (c) continue; ...rest... }
// where the prefix P is idempotent (it never reads a variable it modifies before having written it
// in the same pass), P and c do not depend on the loop counter, and the counter is local to the
// loop, a taken `continue` leaves the state exactly as the next iteration will find it: the next
// iteration recomputes P to the same values, takes the same branch, and so on until the counter
// runs out.  `continue` is then equivalent to `break`, which skips the redundant re-evaluations.
// examples/mandelmesh.frag has this shape (`r = length(z); if (r > 2.0) continue;`): for the 73 % of
// grid corners outside the bailout radius it saves 4 of the 5 length() evaluations.
//
// rotate_guarded_loop (after the two above): a loop that tests an exit right after a cheap prefix,
//     for (init; cond; step) { P; if (c) break; R }
// becomes
//     { init; if (cond) { P; if (c) {} else { for (;;) { R; step; if (cond) {} else { break; } P; if (c) break; } } } }
// -- the same statements in the same order for every path, but everything the compiler hoists out of the
// loop for R's sake (constants, invariant sub-expressions: ~30 instructions for the mandelbulb) now sits
// behind the first test, and a point that leaves at once (96 % of the mandelbulb's grid corners at bounds 5)
// no longer pays for it.  P is duplicated, so it must be plain assignments; R must not `continue` this loop.
//
// pair_sin_cos: within a run of plain statements of one block, sin(e) and cos(e) of a structurally
// identical pure scalar f32 expression e -- with none of e's variables written in between -- become
//     vec2 _sc = sincos_pair(e);   ... _sc.x ... _sc.y ...
// declared before the first of them.  s2m_sincos returns exactly s2m_sin(e) and s2m_cos(e) (one
// argument reduction and one big-argument test instead of two); rotation code and the spherical
// coordinates of examples/mandelmesh.frag have this shape.
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

#include <algorithm>
#include <cstdio>

#include "parse.h"

namespace s2m_frontend {

namespace {

// functions that (transitively) assign module-scope variables: calling one is a side effect
thread_local const std::set<const Function*>* g_state_writers = nullptr;

void reads_of(const Expr& e, std::set<const Var*>& out, bool* impure) {
  if (e.k == Expr::VarRef) out.insert(e.var);
  if (e.k == Expr::UserCall) {
    for (const Var* p : e.fn->params) if (p->by_ref) *impure = true;  // may write through the reference
    if (g_state_writers && g_state_writers->count(e.fn)) *impure = true;
  }
  if (e.k == Expr::AddrOf) *impure = true;
  for (const ExprP& a : e.args) reads_of(*a, out, impure);
}

const Var* assigned_var(const Expr& lhs, bool* partial) {
  const Expr* e = &lhs;
  *partial = false;
  while (e->k == Expr::Swizzle || e->k == Expr::Deref || e->k == Expr::Member || e->k == Expr::Index) {
    if (e->k != Expr::Deref) *partial = true;
    e = e->args[0].get();
  }
  return e->k == Expr::VarRef ? e->var : nullptr;
}

// variables read by the index expressions of an assignment target (a[i].x = ...)
void lhs_index_reads(const Expr& lhs, std::set<const Var*>& out, bool* impure) {
  const Expr* e = &lhs;
  while (e->k == Expr::Swizzle || e->k == Expr::Deref || e->k == Expr::Member || e->k == Expr::Index) {
    if (e->k == Expr::Index) reads_of(*e->args[1], out, impure);
    e = e->args[0].get();
  }
}

bool is_bare_continue(const Stmt& s) {
  if (s.k != Stmt::If || s.else_s || !s.then_s) return false;
  const Stmt* t = s.then_s.get();
  while (t->k == Stmt::Block && t->body.size() == 1) t = t->body[0].get();
  return t->k == Stmt::Continue;
}
void make_break(Stmt& s) {
  Stmt* t = s.then_s.get();
  while (t->k == Stmt::Block && t->body.size() == 1) t = t->body[0].get();
  t->k = Stmt::Break;
}

void writes_anywhere(const Stmt& s, std::set<const Var*>& out) {
  if (s.k == Stmt::Assign) { bool p; if (const Var* v = assigned_var(*s.a, &p)) out.insert(v); }
  if (s.k == Stmt::VarDecl) out.insert(s.var);
  for (const StmtP& c : s.body) writes_anywhere(*c, out);
  if (s.init) writes_anywhere(*s.init, out);
  if (s.cont) writes_anywhere(*s.cont, out);
  if (s.then_s) writes_anywhere(*s.then_s, out);
  if (s.else_s) writes_anywhere(*s.else_s, out);
}

bool try_loop(Stmt& loop) {
  if (loop.k != Stmt::For || loop.body.empty()) return false;
  // the counter(s) must be declared by the for-init, so that they are dead after the loop
  std::set<const Var*> control;  // variables the loop header writes or reads
  if (loop.init) {
    if (loop.init->k != Stmt::VarDecl) return false;
    control.insert(loop.init->var);
  }
  bool impure = false;
  if (loop.cont) { writes_anywhere(*loop.cont, control); if (loop.cont->b) reads_of(*loop.cont->b, control, &impure); }
  if (loop.a) reads_of(*loop.a, control, &impure);
  if (impure) return false;
  {  // whatever the continuing statement advances must be local to the loop: with `break` the
     // counter stops early, which is only unobservable if nobody can read it afterwards
    std::set<const Var*> w;
    if (loop.cont) writes_anywhere(*loop.cont, w);
    for (const Var* v : w)
      if (!loop.init || v != loop.init->var) return false;
  }
  Stmt& body = *loop.body[0];
  bool changed = false;
  std::set<const Var*> all_prefix_writes;
  for (size_t k = 0; k < body.body.size(); ++k) {
    Stmt& st = *body.body[k];
    if (is_bare_continue(st)) {
      // prefix = body.body[0..k)
      std::set<const Var*> w_all;
      bool ok = true;
      for (size_t j = 0; j < k && ok; ++j) {
        const Stmt& p = *body.body[j];
        if (p.k != Stmt::Assign && p.k != Stmt::VarDecl) ok = false;
        else writes_anywhere(p, w_all);
      }
      std::set<const Var*> written;
      for (size_t j = 0; j < k && ok; ++j) {
        const Stmt& p = *body.body[j];
        std::set<const Var*> rd;
        bool imp = false;
        if (p.k == Stmt::VarDecl) { if (p.a) reads_of(*p.a, rd, &imp); }
        else {
          reads_of(*p.b, rd, &imp);
          lhs_index_reads(*p.a, rd, &imp);
          bool partial = false;
          const Var* v = assigned_var(*p.a, &partial);
          if (!v || v->by_ref || v->storage == Var::Global) ok = false;
          if (partial && v) rd.insert(v);
        }
        if (imp) ok = false;
        for (const Var* v : rd) {
          if (control.count(v)) ok = false;                       // depends on the counter
          if (w_all.count(v) && !written.count(v)) ok = false;    // reads last pass's value of something P changes
        }
        if (p.k == Stmt::VarDecl) written.insert(p.var);
        else { bool partial; const Var* v = assigned_var(*p.a, &partial); if (v && !partial) written.insert(v); }
      }
      if (ok) {
        std::set<const Var*> rd;
        bool imp = false;
        reads_of(*st.a, rd, &imp);
        if (imp) ok = false;
        for (const Var* v : rd) {
          if (control.count(v)) ok = false;
          if (w_all.count(v) && !written.count(v)) ok = false;
        }
        for (const Var* v : w_all) if (control.count(v)) ok = false;  // P must not steer the loop header
      }
      if (ok) { make_break(st); changed = true; }
      return changed;  // only the first continue can be rewritten: later ones follow state changes
    }
    if (st.k != Stmt::Assign && st.k != Stmt::VarDecl) return changed;
  }
  return changed;
}

// ---------------------------------------------------------------------------- pair_sin_cos
bool expr_key(const Expr& e, std::string& key, std::set<const Var*>& reads) {  // false: not a pure, nameable expression
  char buf[64];
  switch (e.k) {
    case Expr::Lit:
      snprintf(buf, sizeof buf, "L%d:%d:%a:%lld;", (int)e.ty.sk, e.ty.n, e.lit.f[0], (long long)e.lit.i[0]);
      if (e.ty.n != 1) return false;
      key += buf;
      return true;
    case Expr::VarRef:
      if (e.var->is_ptr || e.var->by_ref) return false;
      reads.insert(e.var);
      snprintf(buf, sizeof buf, "V%d;", e.var->id);
      key += buf;
      return true;
    case Expr::Swizzle:
      key += "S";
      for (int k = 0; k < e.nswz; ++k) key += char('0' + e.swz[k]);
      key += "(";
      if (!expr_key(*e.args[0], key, reads)) return false;
      key += ")";
      return true;
    case Expr::Member:
      key += "M" + std::to_string(e.swz[0]) + "(";
      if (!expr_key(*e.args[0], key, reads)) return false;
      key += ")";
      return true;
    case Expr::Unary: case Expr::Binary: case Expr::Call: case Expr::Construct: case Expr::Convert: case Expr::Ternary: case Expr::Index:
      if (e.k == Expr::Call && (e.callee == "sin" || e.callee == "cos")) return false;  // keeps the hoisted declarations independent
      snprintf(buf, sizeof buf, "E%d:%d:%d:%d:", (int)e.k, (int)e.op, (int)e.ty.sk, e.ty.n + 8 * (int)e.ty.k);
      key += buf;
      key += e.callee;
      key += "(";
      for (const ExprP& a : e.args) { if (!expr_key(*a, key, reads)) return false; key += ","; }
      key += ")";
      return true;
    default: return false;  // UserCall, AddrOf, Deref
  }
}

bool has_user_call(const Expr& e) {
  if (e.k == Expr::UserCall) return true;
  for (const ExprP& a : e.args) if (has_user_call(*a)) return true;
  return false;
}

struct TrigGroup {
  std::vector<Expr*> calls;
  std::set<const Var*> reads;
  size_t first_stmt = 0;
  bool has_sin = false, has_cos = false;
};

struct TrigPairing {
  Module& m;
  int pairs = 0;
  explicit TrigPairing(Module& mod) : m(mod) {}

  bool no_globals = false;  // the statement being scanned calls user functions, which may write globals

  void collect(Expr& e, size_t stmt_index, std::map<std::string, TrigGroup>& groups) {
    for (ExprP& a : e.args) collect(*a, stmt_index, groups);
    if (e.k != Expr::Call || (e.callee != "sin" && e.callee != "cos") || e.args.size() != 1) return;
    if (!(e.ty == Type::scalar(Sk::F32)) || !(e.args[0]->ty == Type::scalar(Sk::F32))) return;
    std::string key;
    std::set<const Var*> reads;
    if (!expr_key(*e.args[0], key, reads)) return;
    if (no_globals)
      for (const Var* v : reads) if (v->storage == Var::Global) return;
    auto it = groups.find(key);
    if (it == groups.end()) {
      TrigGroup g;
      g.reads = reads;
      g.first_stmt = stmt_index;
      it = groups.emplace(key, std::move(g)).first;
    }
    it->second.calls.push_back(&e);
    (e.callee == "sin" ? it->second.has_sin : it->second.has_cos) = true;
  }

  // turns the group's calls into swizzles of one pair variable; returns the declaration to insert
  StmtP materialize(TrigGroup& g) {
    Var* v = m.new_var();
    v->name = "_sc" + std::to_string(pairs++);
    v->ty = Type::vec(Sk::F32, 2);
    v->storage = Var::Local;
    v->immutable = true;
    ExprP call(new Expr());
    call->k = Expr::Call;
    call->callee = "sincos_pair";
    call->ty = v->ty;
    call->args.push_back(g.calls[0]->args[0]);
    call->line = g.calls[0]->line;
    StmtP decl(new Stmt());
    decl->k = Stmt::VarDecl;
    decl->var = v;
    decl->a = call;
    decl->line = call->line;
    for (Expr* e : g.calls) {
      const bool is_sin = e->callee == "sin";
      ExprP ref(new Expr());
      ref->k = Expr::VarRef;
      ref->var = v;
      ref->ty = v->ty;
      e->k = Expr::Swizzle;
      e->callee.clear();
      e->args.assign(1, ref);
      e->nswz = 1;
      e->swz[0] = is_sin ? 0 : 1;
    }
    return decl;
  }

  void block(Stmt& b) {
    std::map<std::string, TrigGroup> groups;
    std::vector<std::pair<size_t, StmtP>> inserts;  // (before statement index, declaration)
    auto finish = [&](TrigGroup& g) { if (g.has_sin && g.has_cos) inserts.emplace_back(g.first_stmt, materialize(g)); };
    auto flush_all = [&] { for (auto& kv : groups) finish(kv.second); groups.clear(); };
    auto invalidate = [&](const Var* v) {
      for (auto it = groups.begin(); it != groups.end();) {
        if (it->second.reads.count(v)) { finish(it->second); it = groups.erase(it); } else ++it;
      }
    };
    for (size_t k = 0; k < b.body.size(); ++k) {
      Stmt& st = *b.body[k];
      const bool plain = st.k == Stmt::VarDecl || st.k == Stmt::Assign || st.k == Stmt::Return || st.k == Stmt::CallStmt;
      if (!plain) { flush_all(); nested(st); continue; }
      std::set<const Var*> rd;
      bool impure = false;
      if (st.a) reads_of(*st.a, rd, &impure);
      if (st.b) reads_of(*st.b, rd, &impure);
      if (impure) { flush_all(); continue; }  // a by-reference user call may write anything it is handed
      no_globals = (st.a && has_user_call(*st.a)) || (st.b && has_user_call(*st.b));
      if (no_globals) {  // ... and any call may write globals
        std::set<const Var*> globals;
        for (auto& kv : groups) for (const Var* v : kv.second.reads) if (v->storage == Var::Global) globals.insert(v);
        for (const Var* v : globals) invalidate(v);
      }
      if (st.k == Stmt::Assign) {
        collect(*st.b, k, groups);
        collect(*st.a, k, groups);  // index expressions on the left are reads too
        bool partial;
        if (const Var* v = assigned_var(*st.a, &partial)) invalidate(v); else flush_all();
      } else {
        if (st.a) collect(*st.a, k, groups);
        if (st.k == Stmt::VarDecl) invalidate(st.var);
      }
    }
    flush_all();
    std::stable_sort(inserts.begin(), inserts.end(), [](const std::pair<size_t, StmtP>& x, const std::pair<size_t, StmtP>& y) { return x.first > y.first; });
    for (auto& ins : inserts) b.body.insert(b.body.begin() + (long)ins.first, ins.second);
  }

  void nested(Stmt& s) {
    if (s.k == Stmt::Block) { block(s); return; }
    for (StmtP& c : s.body) nested(*c);
    if (s.then_s) nested(*s.then_s);
    if (s.else_s) nested(*s.else_s);
    // for-init / continuing statements and loop conditions are left alone: they are re-evaluated
  }
};

// ---------------------------------------------------------------------------- rotate_guarded_loop
bool is_bare_break(const Stmt& s) {
  if (s.k != Stmt::If || s.else_s || !s.then_s) return false;
  const Stmt* t = s.then_s.get();
  while (t->k == Stmt::Block && t->body.size() == 1) t = t->body[0].get();
  return t->k == Stmt::Break;
}
// a `continue` that belongs to the loop whose body this is (nested loops keep theirs)
bool continues_this_loop(const Stmt& s) {
  if (s.k == Stmt::Continue) return true;
  if (s.k == Stmt::For || s.k == Stmt::While || s.k == Stmt::DoWhile || s.k == Stmt::Loop) return false;
  for (const StmtP& c : s.body) if (continues_this_loop(*c)) return true;
  if (s.then_s && continues_this_loop(*s.then_s)) return true;
  if (s.else_s && continues_this_loop(*s.else_s)) return true;
  return false;
}
StmtP new_stmt(Stmt::K k, int line) { StmtP s(new Stmt()); s->k = k; s->line = line; return s; }

bool rotate_loop(Stmt& loop) {
  if (loop.k != Stmt::For || loop.body.size() != 1 || loop.body[0]->k != Stmt::Block) return false;
  if (loop.init && loop.init->k != Stmt::VarDecl) return false;
  if (loop.cont && loop.cont->k != Stmt::Assign && loop.cont->k != Stmt::CallStmt) return false;
  std::vector<StmtP>& b = loop.body[0]->body;
  size_t k = 0;
  while (k < b.size() && b[k]->k == Stmt::Assign) ++k;      // P: plain assignments only (it is emitted twice)
  if (k == 0 || k >= b.size() || !is_bare_break(*b[k]) || k + 1 >= b.size()) return false;
  for (size_t j = k + 1; j < b.size(); ++j) if (continues_this_loop(*b[j])) return false;
  const int line = loop.line;
  const std::vector<StmtP> prefix(b.begin(), b.begin() + (long)k);
  const StmtP exit_test = b[k];
  const std::vector<StmtP> rest(b.begin() + (long)k + 1, b.end());
  // for (;;) { R; step; if (cond) {} else { break; } P; if (c) break; }
  StmtP inner_body = new_stmt(Stmt::Block, line);
  inner_body->body = rest;
  if (loop.cont) inner_body->body.push_back(loop.cont);
  if (loop.a) {
    StmtP leave = new_stmt(Stmt::If, line);
    leave->a = loop.a;
    leave->then_s = new_stmt(Stmt::Block, line);
    leave->else_s = new_stmt(Stmt::Block, line);
    leave->else_s->body.push_back(new_stmt(Stmt::Break, line));
    inner_body->body.push_back(leave);
  }
  inner_body->body.insert(inner_body->body.end(), prefix.begin(), prefix.end());
  inner_body->body.push_back(exit_test);
  StmtP inner = new_stmt(Stmt::For, line);
  inner->body.push_back(inner_body);
  // P; if (c) {} else { inner }
  StmtP first_test = new_stmt(Stmt::If, exit_test->line);
  first_test->a = exit_test->a;
  first_test->then_s = new_stmt(Stmt::Block, line);
  first_test->else_s = new_stmt(Stmt::Block, line);
  first_test->else_s->body.push_back(inner);
  StmtP entered = new_stmt(Stmt::Block, line);
  entered->body = prefix;
  entered->body.push_back(first_test);
  std::vector<StmtP> out;
  if (loop.init) out.push_back(loop.init);
  if (loop.a) {
    StmtP guard = new_stmt(Stmt::If, line);
    guard->a = loop.a;
    guard->then_s = entered;
    out.push_back(guard);
  } else {
    out.push_back(entered);
  }
  loop.k = Stmt::Block;
  loop.body = out;
  loop.init.reset(); loop.cont.reset(); loop.a.reset();
  return true;
}
void rotate_walk(Stmt& s, int* count) {
  for (StmtP& c : s.body) rotate_walk(*c, count);
  if (s.then_s) rotate_walk(*s.then_s, count);
  if (s.else_s) rotate_walk(*s.else_s, count);
  if (rotate_loop(s)) ++*count;
}

void walk(Stmt& s, int* count) {
  for (StmtP& c : s.body) walk(*c, count);
  if (s.then_s) walk(*s.then_s, count);
  if (s.else_s) walk(*s.else_s, count);
  if (try_loop(s)) ++*count;
}

}  // namespace

void callees_of(const Expr& e, std::set<const Function*>& out) {
  if (e.k == Expr::UserCall) out.insert(e.fn);
  for (const ExprP& a : e.args) callees_of(*a, out);
}
void callees_of(const Stmt& s, std::set<const Function*>& out) {
  if (s.a) callees_of(*s.a, out);
  if (s.b) callees_of(*s.b, out);
  if (s.break_if) callees_of(*s.break_if, out);
  for (const StmtP& c : s.body) callees_of(*c, out);
  if (s.init) callees_of(*s.init, out);
  if (s.cont) callees_of(*s.cont, out);
  if (s.then_s) callees_of(*s.then_s, out);
  if (s.else_s) callees_of(*s.else_s, out);
}

std::set<const Function*> state_writers(const Module& m) {
  std::set<const Function*> w;
  std::map<const Function*, std::set<const Function*>> calls;
  for (const auto& f : m.functions) {
    if (!f->body) continue;
    std::set<const Var*> written;
    writes_anywhere(*f->body, written);
    for (const Var* v : written) if (v->storage == Var::Global) w.insert(f.get());
    callees_of(*f->body, calls[f.get()]);
  }
  for (bool changed = true; changed;) {
    changed = false;
    for (const auto& kv : calls)
      if (!w.count(kv.first))
        for (const Function* c : kv.second)
          if (w.count(c)) { w.insert(kv.first); changed = true; break; }
  }
  return w;
}

namespace {
void eager_conditionals_in(const Expr& e, int line) {
  if (e.k == Expr::Ternary) {
    for (int arm = 1; arm <= 2; ++arm) {
      std::set<const Var*> reads;
      bool impure = false;
      reads_of(*e.args[(size_t)arm], reads, &impure);
      if (impure)
        throw FrontendError(11 /*S2M_ERR_UNSUPPORTED*/, "unsupported near line " + std::to_string(line) +
                                ": an arm of ?: calls a function with a side effect (it assigns a module-scope variable or has an out / inout "
                                "parameter) in a position where no statement can be issued for it (an `else if` condition, the right side of "
                                "&& or ||, a nested ?:); GLSL evaluates only the chosen arm -- assign the result to a variable first");
    }
  }
  for (const ExprP& a : e.args) if (a) eager_conditionals_in(*a, line);
}
void eager_conditionals_in(const Stmt& s, int line) {
  if (s.line > 0) line = s.line;
  for (const ExprP* e : {&s.a, &s.b, &s.break_if}) if (*e) eager_conditionals_in(**e, line);
  for (const StmtP& c : s.body) eager_conditionals_in(*c, line);
  for (const StmtP* c : {&s.init, &s.cont, &s.then_s, &s.else_s}) if (*c) eager_conditionals_in(**c, line);
}
}  // namespace

// GLSL's `c ? t : f` that stayed an expression (parse_glsl.cpp parse_conditional_arms) becomes WGSL select(),
// which evaluates both arms: harmless for pure arms, wrong for an arm whose call has a side effect.
void check_eager_conditionals(const Module& m) {
  const std::set<const Function*> writers = state_writers(m);
  g_state_writers = &writers;
  struct Reset { ~Reset() { g_state_writers = nullptr; } } reset;
  for (const auto& f : m.functions)
    if (f->body) eager_conditionals_in(*f->body, f->line);
}

int optimize_module(Module& m) {
  int count = 0;
  const std::set<const Function*> writers = state_writers(m);
  g_state_writers = &writers;
  struct Reset { ~Reset() { g_state_writers = nullptr; } } reset;
  TrigPairing pairing(m);
  for (auto& f : m.functions) {
    if (!f->body) continue;
    walk(*f->body, &count);
    pairing.nested(*f->body);
    if (!getenv("S2M_NO_LOOP_ROTATION")) rotate_walk(*f->body, &count);   // last: it emits a loop's prefix twice
  }
  return count + pairing.pairs;
}

}  // namespace s2m_frontend
