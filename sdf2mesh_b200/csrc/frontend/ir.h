// ir.h -- naga-shaped typed IR shared by the WGSL and GLSL front-ends and the CUDA / WGSL writers.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace s2m_frontend {

enum class Sk : uint8_t { Bool, I32, U32, F32, AInt, AFloat };  // A* = WGSL abstract numerics

struct Type {
  enum K : uint8_t { Void, Scalar, Vector, Matrix } k = Void;  // Matrix: n x n of f32, column-major
  Sk sk = Sk::F32;
  int n = 1;  // vector width
  static Type scalar(Sk s) { Type t; t.k = Scalar; t.sk = s; t.n = 1; return t; }
  static Type vec(Sk s, int n) { Type t; t.k = n == 1 ? Scalar : Vector; t.sk = s; t.n = n; return t; }
  static Type mat(int n) { Type t; t.k = Matrix; t.sk = Sk::F32; t.n = n; return t; }
  static Type void_() { return Type(); }
  bool is_void() const { return k == Void; }
  bool is_scalar() const { return k == Scalar; }
  bool is_vector() const { return k == Vector; }
  bool is_matrix() const { return k == Matrix; }
  bool is_abstract() const { return !is_void() && (sk == Sk::AInt || sk == Sk::AFloat); }
  bool is_float() const { return !is_void() && (sk == Sk::F32 || sk == Sk::AFloat); }
  bool is_int() const { return !is_void() && (sk == Sk::I32 || sk == Sk::U32 || sk == Sk::AInt); }
  bool is_bool() const { return !is_void() && sk == Sk::Bool; }
  bool operator==(const Type& o) const { return k == o.k && (k == Void || (sk == o.sk && n == o.n)); }
  bool operator!=(const Type& o) const { return !(*this == o); }
  Type with_sk(Sk s) const { Type t = *this; t.sk = s; return t; }
  std::string str() const;
};

struct ConstVal {  // value of a constant expression; component c in f[c] (floats) or i[c] (ints, bools)
  Type ty;
  double f[4] = {0, 0, 0, 0};
  int64_t i[4] = {0, 0, 0, 0};
};

struct Function;
struct Var {
  std::string name;
  Type ty;
  enum Storage : uint8_t { Local, Param, Global, ModuleConst } storage = Local;
  bool immutable = false;   // let / const / `in`-less const param
  bool by_ref = false;      // GLSL out/inout, WGSL ptr<function,T>
  bool is_ptr = false;      // WGSL pointer parameter (needs explicit * to access)
  bool has_const = false;   // value known at compile time
  ConstVal cval;
  bool written = false;     // assigned somewhere (globals)
  int id = 0;
};

enum class Op : uint8_t {
  Add, Sub, Mul, Div, Rem, Neg, Not, BitNot, And, Or, BitAnd, BitOr, BitXor, Shl, Shr,
  Lt, Le, Gt, Ge, Eq, Ne
};

struct Expr;
typedef std::shared_ptr<Expr> ExprP;
struct Expr {
  enum K : uint8_t { Lit, VarRef, Unary, Binary, Call, UserCall, Construct, Swizzle, Ternary, Convert, AddrOf, Deref } k = Lit;
  Type ty;
  ConstVal lit;            // Lit
  Var* var = nullptr;      // VarRef
  Op op = Op::Add;         // Unary / Binary
  std::string callee;      // Call: canonical builtin name
  Function* fn = nullptr;  // UserCall
  std::vector<ExprP> args; // operands / call args / constructor args / swizzle base / ternary (c,t,f)
  int swz[4] = {0, 0, 0, 0};
  int nswz = 0;
  int line = 0;
};

struct Stmt;
typedef std::shared_ptr<Stmt> StmtP;
struct Stmt {
  enum K : uint8_t { Block, VarDecl, Assign, If, For, While, DoWhile, Loop, Break, Continue, Return, CallStmt, Discard } k = Block;
  std::vector<StmtP> body;       // Block / loop bodies (body[0] for If-then? see below)
  Var* var = nullptr;            // VarDecl
  ExprP a, b;                    // VarDecl: a=init; Assign: a=lhs, b=rhs; If/While/DoWhile: a=cond; Return: a=value; CallStmt: a
  StmtP init, cont;              // For: init, continuing;  Loop: cont = continuing block
  StmtP then_s, else_s;          // If
  ExprP break_if;                // Loop: `break if` in continuing
  int line = 0;
};

struct Function {
  std::string name;
  Type ret;
  std::vector<Var*> params;
  StmtP body;        // null for prototypes
  bool builtin_lib = false;  // provided by s2m_sdf3d_lib.h / kernels_jit.cuh; not emitted
  bool is_entry = false;     // GLSL main / WGSL @fragment/@compute etc.
  int line = 0;
};

struct Module {
  std::vector<std::unique_ptr<Var>> vars;        // owns all Vars
  std::vector<std::unique_ptr<Function>> functions;
  std::vector<Var*> globals;                     // in declaration order (Global + ModuleConst)
  std::map<const Var*, ExprP> global_init;       // initializer expressions of globals
  Var* new_var() { vars.emplace_back(new Var()); vars.back()->id = (int)vars.size(); return vars.back().get(); }
};

struct FrontendError : std::runtime_error {
  int status;
  FrontendError(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

}  // namespace s2m_frontend
