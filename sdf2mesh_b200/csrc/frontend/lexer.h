// lexer.h -- tokenizer shared by the WGSL and GLSL parsers.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "ir.h"

namespace s2m_frontend {

struct Token {
  enum K { End, Ident, Int, Float, Punct, Directive } k = End;  // Directive: a GLSL '#...' line (text = the line)
  std::string text;     // identifier / punctuation / literal spelling
  double fval = 0;
  int64_t ival = 0;
  char suffix = 0;      // 'f', 'u', 'i', 'h' or 0
  int line = 1, col = 1;
};

struct LexOptions {
  bool glsl = false;  // GLSL: preprocessor lines, no nested block comments
};

// Object-like and function-like #define macros (GLSL only).
struct Macro {
  bool function_like = false;
  std::vector<std::string> params;
  std::vector<Token> body;
};

class Lexer {
 public:
  Lexer(const std::string& src, const LexOptions& opt) : s_(normalized(src)), opt_(opt) {}
  // Editors on other platforms leave a UTF-8 byte-order mark and \r\n (or bare \r) line ends; neither means
  // anything to WGSL or GLSL, and `\` + `\r\n` must still continue a preprocessor line.
  static std::string normalized(const std::string& src) {
    std::string out;
    out.reserve(src.size());
    size_t i = src.compare(0, 3, "\xef\xbb\xbf") == 0 ? 3 : 0;
    for (; i < src.size(); ++i) {
      if (src[i] == '\r') { out += '\n'; if (i + 1 < src.size() && src[i + 1] == '\n') ++i; }
      else out += src[i];
    }
    return out;
  }
  std::vector<Token> run() {
    std::vector<Token> raw;
    for (;;) {
      Token t = next();
      raw.push_back(t);
      if (t.k == Token::End) break;
    }
    if (!opt_.glsl) return raw;
    return preprocess(raw);
  }

 private:
  const std::string s_;   // normalized copy: no byte-order mark, \n line ends
  LexOptions opt_;
  size_t i_ = 0;
  int line_ = 1, col_ = 1;
  std::map<std::string, Macro> macros_;
  bool at_line_start_ = true;

  [[noreturn]] void err(const std::string& m, int line, int col) {
    throw FrontendError(3 /*S2M_ERR_PARSE*/, "parse error at " + std::to_string(line) + ":" + std::to_string(col) + ": " + m);
  }
  char peek(size_t o = 0) const { return i_ + o < s_.size() ? s_[i_ + o] : '\0'; }
  void adv() {
    if (s_[i_] == '\n') { ++line_; col_ = 1; at_line_start_ = true; }
    else { ++col_; if (!isspace((unsigned char)s_[i_])) at_line_start_ = false; }
    ++i_;
  }
  void skip_ws_and_comments() {
    for (;;) {
      while (i_ < s_.size() && isspace((unsigned char)peek())) adv();
      if (peek() == '/' && peek(1) == '/') {
        while (i_ < s_.size() && peek() != '\n') adv();
        continue;
      }
      if (peek() == '/' && peek(1) == '*') {
        int depth = 0, l = line_, c = col_;
        do {
          if (peek() == '/' && peek(1) == '*') { ++depth; adv(); adv(); }
          else if (peek() == '*' && peek(1) == '/') { --depth; adv(); adv(); if (opt_.glsl) depth = 0; }
          else if (i_ >= s_.size()) err("unterminated block comment", l, c);
          else adv();
        } while (depth > 0);
        continue;
      }
      break;
    }
  }
  // reads a '#...' line (with backslash continuations) into one Directive token
  Token directive() {
    Token t;
    t.k = Token::Directive; t.line = line_; t.col = col_;
    while (i_ < s_.size() && peek() != '\n') {
      if (peek() == '\\' && peek(1) == '\n') { adv(); adv(); t.text += ' '; continue; }
      if (peek() == '/' && peek(1) == '/') { while (i_ < s_.size() && peek() != '\n') adv(); break; }
      t.text += peek();
      adv();
    }
    return t;
  }

  static std::string directive_name(const std::string& text, size_t* rest) {
    size_t p = 1;
    while (p < text.size() && isspace((unsigned char)text[p])) ++p;
    size_t q = p;
    while (q < text.size() && isalpha((unsigned char)text[q])) ++q;
    *rest = q;
    return text.substr(p, q - p);
  }

  void define_macro(const Token& d, size_t q) {
    const std::string& text = d.text;
    while (q < text.size() && isspace((unsigned char)text[q])) ++q;
    size_t r = q;
    while (r < text.size() && (isalnum((unsigned char)text[r]) || text[r] == '_')) ++r;
    const std::string mname = text.substr(q, r - q);
    if (mname.empty()) err("#define without a name", d.line, d.col);
    Macro m;
    if (r < text.size() && text[r] == '(') {
      m.function_like = true;
      ++r;
      std::string cur;
      for (; r < text.size() && text[r] != ')'; ++r) {
        if (text[r] == ',') { m.params.push_back(cur); cur.clear(); }
        else if (!isspace((unsigned char)text[r])) cur += text[r];
      }
      if (!cur.empty()) m.params.push_back(cur);
      if (r < text.size()) ++r;
    }
    for (Token& t : lex_fragment(text.substr(r), d)) m.body.push_back(t);
    macros_[mname] = m;
  }

  std::vector<Token> lex_fragment(const std::string& body, const Token& at) {
    LexOptions o; o.glsl = false;
    Lexer sub(body, o);
    std::vector<Token> out;
    for (;;) { Token t = sub.next(); if (t.k == Token::End) break; t.line = at.line; t.col = at.col; out.push_back(t); }
    return out;
  }

  // ---- #if expression: integers, defined(X), ! - + * / % < > <= >= == != && || and parentheses
  struct IfEval {
    const std::vector<Token>& t;
    size_t i = 0;
    Lexer* lx;
    const Token& at;
    bool is(const char* p) const { return i < t.size() && t[i].k == Token::Punct && t[i].text == p; }
    long long primary() {
      if (i >= t.size()) lx->err("malformed #if expression", at.line, at.col);
      if (is("(")) { ++i; long long v = lor(); if (!is(")")) lx->err("expected ')' in #if", at.line, at.col); ++i; return v; }
      if (is("!")) { ++i; return !primary(); }
      if (is("-")) { ++i; return -primary(); }
      if (is("+")) { ++i; return primary(); }
      const Token& k = t[i++];
      if (k.k == Token::Int) return k.ival;
      if (k.k == Token::Float) return (long long)k.fval;
      if (k.k == Token::Ident) return 0;  // unknown identifiers evaluate to 0, as in cpp
      lx->err("unexpected token '" + k.text + "' in #if", at.line, at.col);
    }
    long long mul() { long long v = primary(); while (is("*") || is("/") || is("%")) { const char o = t[i++].text[0]; long long r = primary(); v = o == '*' ? v * r : (r == 0 ? 0 : (o == '/' ? v / r : v % r)); } return v; }
    long long add() { long long v = mul(); while (is("+") || is("-")) { const char o = t[i++].text[0]; long long r = mul(); v = o == '+' ? v + r : v - r; } return v; }
    long long rel() { long long v = add(); while (is("<") || is(">") || is("<=") || is(">=")) { const std::string o = t[i++].text; long long r = add(); v = o == "<" ? v < r : o == ">" ? v > r : o == "<=" ? v <= r : v >= r; } return v; }
    long long eq() { long long v = rel(); while (is("==") || is("!=")) { const bool e = t[i++].text == "=="; long long r = rel(); v = e ? v == r : v != r; } return v; }
    long long land() { long long v = eq(); while (is("&&")) { ++i; long long r = eq(); v = v && r; } return v; }
    long long lor() { long long v = land(); while (is("||")) { ++i; long long r = land(); v = v || r; } return v; }
  };
  bool eval_if(const Token& d, size_t q) {
    std::vector<Token> toks = lex_fragment(d.text.substr(q), d);
    // defined(X) / defined X  before macro expansion
    std::vector<Token> pre;
    for (size_t k = 0; k < toks.size(); ++k) {
      if (toks[k].k == Token::Ident && toks[k].text == "defined") {
        size_t j = k + 1;
        const bool paren = j < toks.size() && toks[j].text == "(";
        if (paren) ++j;
        if (j >= toks.size() || toks[j].k != Token::Ident) err("malformed defined()", d.line, d.col);
        Token v; v.k = Token::Int; v.ival = macros_.count(toks[j].text) ? 1 : 0; v.line = d.line; v.col = d.col;
        pre.push_back(v);
        k = paren ? j + 1 : j;
        continue;
      }
      pre.push_back(toks[k]);
    }
    std::vector<Token> ex;
    expand(pre, ex, 0);
    IfEval ev{ex, 0, this, d};
    const long long v = ev.lor();
    if (ev.i != ex.size()) err("trailing tokens in #if expression", d.line, d.col);
    return v != 0;
  }

  // Ordered pass over the raw tokens: conditionals, #define / #undef, macro expansion.
  std::vector<Token> preprocess(const std::vector<Token>& raw) {
    struct Cond { bool parent, taken, active; };
    std::vector<Cond> stack;
    std::vector<Token> out, pending;
    auto active = [&] { return stack.empty() || stack.back().active; };
    auto flush = [&] { if (!pending.empty()) { expand(pending, out, 0); pending.clear(); } };
    for (const Token& t : raw) {
      if (t.k != Token::Directive) {
        if (t.k == Token::End) { flush(); out.push_back(t); break; }
        if (active()) pending.push_back(t);
        continue;
      }
      size_t q = 0;
      const std::string name = directive_name(t.text, &q);
      if (name == "ifdef" || name == "ifndef" || name == "if") {
        flush();
        bool v = false;
        if (active()) {
          if (name == "if") v = eval_if(t, q);
          else {
            std::vector<Token> id = lex_fragment(t.text.substr(q), t);
            if (id.empty() || id[0].k != Token::Ident) err("#" + name + " needs an identifier", t.line, t.col);
            v = macros_.count(id[0].text) > 0;
            if (name == "ifndef") v = !v;
          }
        }
        const bool par = active();
        stack.push_back({par, par && v, par && v});
      } else if (name == "elif" || name == "else") {
        flush();
        if (stack.empty()) err("#" + name + " without #if", t.line, t.col);
        Cond& c = stack.back();
        bool v = name == "else" ? true : (c.parent && !c.taken ? eval_if(t, q) : false);
        c.active = c.parent && !c.taken && v;
        if (c.active) c.taken = true;
      } else if (name == "endif") {
        flush();
        if (stack.empty()) err("#endif without #if", t.line, t.col);
        stack.pop_back();
      } else if (!active()) {
        continue;
      } else if (name == "define") {
        flush();
        define_macro(t, q);
      } else if (name == "undef") {
        flush();
        std::vector<Token> id = lex_fragment(t.text.substr(q), t);
        if (!id.empty()) macros_.erase(id[0].text);
      } else if (name == "include") {
        err("#include is not supported in GLSL input", t.line, t.col);
      } else if (name == "error") {
        err("#error" + t.text.substr(q), t.line, t.col);
      }
      // #version, #extension, #pragma, #line: ignored
    }
    if (!stack.empty()) err("unterminated #if", raw.back().line, raw.back().col);
    return out;
  }

  void expand(const std::vector<Token>& in, std::vector<Token>& out, int depth) {
    if (depth > 32) err("macro expansion too deep", in.empty() ? 0 : in[0].line, 0);
    for (size_t k = 0; k < in.size(); ++k) {
      const Token& t = in[k];
      auto it = t.k == Token::Ident ? macros_.find(t.text) : macros_.end();
      if (it == macros_.end()) { out.push_back(t); continue; }
      const Macro& m = it->second;
      if (!m.function_like) {
        std::vector<Token> body = m.body;
        for (auto& b : body) { b.line = t.line; b.col = t.col; }
        expand(body, out, depth + 1);
        continue;
      }
      if (k + 1 >= in.size() || in[k + 1].text != "(") { out.push_back(t); continue; }
      std::vector<std::vector<Token>> args(1);
      size_t j = k + 2;
      int par = 1;
      for (; j < in.size(); ++j) {
        if (in[j].k == Token::End) err("unterminated macro call " + t.text, t.line, t.col);
        if (in[j].k == Token::Punct && in[j].text == "(") ++par;
        if (in[j].k == Token::Punct && in[j].text == ")") { if (--par == 0) break; }
        if (par == 1 && in[j].k == Token::Punct && in[j].text == ",") { args.emplace_back(); continue; }
        args.back().push_back(in[j]);
      }
      if (args.size() == 1 && args[0].empty() && m.params.empty()) args.clear();
      if (args.size() != m.params.size()) err("macro " + t.text + " expects " + std::to_string(m.params.size()) + " arguments", t.line, t.col);
      std::vector<Token> body;
      for (const Token& b : m.body) {
        bool sub = false;
        if (b.k == Token::Ident)
          for (size_t a = 0; a < m.params.size(); ++a)
            if (m.params[a] == b.text) { body.insert(body.end(), args[a].begin(), args[a].end()); sub = true; break; }
        if (!sub) { Token c = b; c.line = t.line; c.col = t.col; body.push_back(c); }
      }
      expand(body, out, depth + 1);
      k = j;
    }
  }

 public:
  Token next() {
    skip_ws_and_comments();
    Token t;
    t.line = line_; t.col = col_;
    if (i_ >= s_.size()) { t.k = Token::End; return t; }
    if (opt_.glsl && peek() == '#' && at_line_start_) return directive();
    const char c = peek();
    if (isalpha((unsigned char)c) || c == '_') {
      while (isalnum((unsigned char)peek()) || peek() == '_') { t.text += peek(); adv(); }
      t.k = Token::Ident;
      return t;
    }
    if (isdigit((unsigned char)c) || (c == '.' && isdigit((unsigned char)peek(1)))) return number(t);
    static const char* three[] = {"<<=", ">>="};
    static const char* two[] = {"->", "==", "!=", "<=", ">=", "&&", "||", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "++", "--", "<<", ">>"};
    for (const char* p : three)
      if (s_.compare(i_, 3, p) == 0) { t.k = Token::Punct; t.text = p; adv(); adv(); adv(); return t; }
    for (const char* p : two)
      if (s_.compare(i_, 2, p) == 0) { t.k = Token::Punct; t.text = p; adv(); adv(); return t; }
    t.k = Token::Punct;
    t.text = std::string(1, c);
    adv();
    return t;
  }

 private:
  Token number(Token t) {
    std::string sp;
    bool is_float = false, hex = false;
    if (peek() == '0' && (peek(1) == 'x' || peek(1) == 'X')) {
      hex = true;
      sp += peek(); adv(); sp += peek(); adv();
      while (isxdigit((unsigned char)peek())) { sp += peek(); adv(); }
    } else {
      while (isdigit((unsigned char)peek())) { sp += peek(); adv(); }
      if (peek() == '.' && !(isalpha((unsigned char)peek(1)) && peek(1) != 'e' && peek(1) != 'E' && peek(1) != 'f' && peek(1) != 'F' && peek(1) != 'h')) {
        is_float = true; sp += peek(); adv();
        while (isdigit((unsigned char)peek())) { sp += peek(); adv(); }
      }
      if ((peek() == 'e' || peek() == 'E') && (isdigit((unsigned char)peek(1)) || ((peek(1) == '+' || peek(1) == '-') && isdigit((unsigned char)peek(2))))) {
        is_float = true; sp += peek(); adv();
        if (peek() == '+' || peek() == '-') { sp += peek(); adv(); }
        while (isdigit((unsigned char)peek())) { sp += peek(); adv(); }
      }
    }
    char suf = 0;
    if (peek() == 'f' || peek() == 'F' || peek() == 'h') { suf = (char)tolower(peek()); adv(); if (!hex) is_float = true; }
    else if (peek() == 'u' || peek() == 'U') { suf = 'u'; adv(); }
    else if (peek() == 'i') { suf = 'i'; adv(); }
    if (isalnum((unsigned char)peek()) || peek() == '_') err("malformed number '" + sp + std::string(1, peek()) + "'", t.line, t.col);
    t.text = sp;
    t.suffix = suf;
    if (suf == 'h') err("f16 literals are not supported", t.line, t.col);
    if (is_float) { t.k = Token::Float; t.fval = strtod(sp.c_str(), nullptr); }
    else { t.k = Token::Int; t.ival = (int64_t)strtoull(sp.c_str(), nullptr, hex ? 16 : 10); t.fval = (double)t.ival; }
    return t;
  }
};

}  // namespace s2m_frontend
