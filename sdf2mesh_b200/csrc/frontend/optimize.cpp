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
#include <set>

#include "parse.h"

namespace s2m_frontend {

namespace {

void reads_of(const Expr& e, std::set<const Var*>& out, bool* impure) {
  if (e.k == Expr::VarRef) out.insert(e.var);
  if (e.k == Expr::UserCall) {
    for (const Var* p : e.fn->params) if (p->by_ref) *impure = true;  // may write through the reference
  }
  if (e.k == Expr::AddrOf) *impure = true;
  for (const ExprP& a : e.args) reads_of(*a, out, impure);
}

const Var* assigned_var(const Expr& lhs, bool* partial) {
  const Expr* e = &lhs;
  *partial = false;
  while (e->k == Expr::Swizzle || e->k == Expr::Deref) {
    if (e->k == Expr::Swizzle) *partial = true;
    e = e->args[0].get();
  }
  return e->k == Expr::VarRef ? e->var : nullptr;
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

void walk(Stmt& s, int* count) {
  for (StmtP& c : s.body) walk(*c, count);
  if (s.then_s) walk(*s.then_s, count);
  if (s.else_s) walk(*s.else_s, count);
  if (try_loop(s)) ++*count;
}

}  // namespace

int optimize_module(Module& m) {
  int count = 0;
  for (auto& f : m.functions)
    if (f->body) walk(*f->body, &count);
  return count;
}

}  // namespace s2m_frontend
