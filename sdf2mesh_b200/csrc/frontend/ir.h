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

struct StructDef;
struct ArrayDef;

struct Type {
  enum K : uint8_t { Void, Scalar, Vector, Matrix, Struct, Array } k = Void;  // Matrix: n x n of f32, column-major
  Sk sk = Sk::F32;
  int n = 1;  // vector width
  const StructDef* sdef = nullptr;  // Struct: definition owned by the Module
  const ArrayDef* adef = nullptr;   // Array: interned (element type, length) owned by the Module
  static Type scalar(Sk s) { Type t; t.k = Scalar; t.sk = s; t.n = 1; return t; }
  static Type vec(Sk s, int n) { Type t; t.k = n == 1 ? Scalar : Vector; t.sk = s; t.n = n; return t; }
  static Type mat(int n) { Type t; t.k = Matrix; t.sk = Sk::F32; t.n = n; return t; }
  static Type void_() { return Type(); }
  static Type struct_(const StructDef* d) { Type t; t.k = Struct; t.sdef = d; return t; }
  static Type array_(const ArrayDef* d) { Type t; t.k = Array; t.adef = d; return t; }
  bool is_struct() const { return k == Struct; }
  bool is_array() const { return k == Array; }
  bool is_aggregate() const { return k == Struct || k == Array; }
  bool is_void() const { return k == Void; }
  bool is_scalar() const { return k == Scalar; }
  bool is_vector() const { return k == Vector; }
  bool is_matrix() const { return k == Matrix; }
  bool is_numeric_kind() const { return k == Scalar || k == Vector || k == Matrix; }
  bool is_abstract() const { return is_numeric_kind() && (sk == Sk::AInt || sk == Sk::AFloat); }
  bool is_float() const { return is_numeric_kind() && (sk == Sk::F32 || sk == Sk::AFloat); }
  bool is_int() const { return is_numeric_kind() && (sk == Sk::I32 || sk == Sk::U32 || sk == Sk::AInt); }
  bool is_bool() const { return is_numeric_kind() && sk == Sk::Bool; }
  bool operator==(const Type& o) const {
    if (k != o.k) return false;
    if (k == Void) return true;
    if (k == Struct) return sdef == o.sdef;
    if (k == Array) return adef == o.adef;
    return sk == o.sk && n == o.n;
  }
  bool operator!=(const Type& o) const { return !(*this == o); }
  Type with_sk(Sk s) const { Type t = *this; t.sk = s; return t; }
  std::string str() const;
};

struct StructDef {
  std::string name;
  std::vector<std::string> field_names;
  std::vector<Type> field_types;
  int field(const std::string& f) const {
    for (size_t i = 0; i < field_names.size(); ++i) if (field_names[i] == f) return (int)i;
    return -1;
  }
};
struct ArrayDef {
  Type elem;
  int len = 0;
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
  enum K : uint8_t { Lit, VarRef, Unary, Binary, Call, UserCall, Construct, Swizzle, Ternary, Convert, AddrOf, Deref,
           Member /* args[0].field swz[0] of a struct */, Index /* args[0][args[1]]: array element, vector component, matrix column */ } k = Lit;
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
  enum K : uint8_t { Block, VarDecl, Assign, If, For, While, DoWhile, Loop, Break, Continue, Return, CallStmt, Discard, Switch, Case } k = Block;
  std::vector<StmtP> body;       // Block / loop bodies (body[0] for If-then? see below)
  Var* var = nullptr;            // VarDecl
  ExprP a, b;                    // VarDecl: a=init; Assign: a=lhs, b=rhs; If/While/DoWhile: a=cond; Return: a=value; CallStmt: a
  StmtP init, cont;              // For: init, continuing;  Loop: cont = continuing block
  StmtP then_s, else_s;          // If
  ExprP break_if;                // Loop: `break if` in continuing
  std::vector<int64_t> case_values;  // Case: selector values (Switch: a = selector, body = Case statements, each with body[0] = Block;
  bool is_default = false;           //       no fall-through: a case ends where the next begins; Break leaves the switch)
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
  std::vector<std::unique_ptr<StructDef>> structs;   // in declaration order
  std::vector<std::unique_ptr<ArrayDef>> arrays;     // interned
  // GLSL functions left out because their bodies need something this engine has no counterpart for
  // (texture sampling, screen-space derivatives, ...): name and reason.  Only calling one is an error.
  std::vector<std::pair<std::string, std::string>> dropped;
  Var* new_var() { vars.emplace_back(new Var()); vars.back()->id = (int)vars.size(); return vars.back().get(); }
  Type array_of(const Type& elem, int len) {
    for (const auto& a : arrays) if (a->elem == elem && a->len == len) return Type::array_(a.get());
    arrays.emplace_back(new ArrayDef());
    arrays.back()->elem = elem;
    arrays.back()->len = len;
    return Type::array_(arrays.back().get());
  }
};

struct FrontendError : std::runtime_error {
  int status;
  FrontendError(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

}  // namespace s2m_frontend
