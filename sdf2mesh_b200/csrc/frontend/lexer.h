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
  enum K { End, Ident, Int, Float, Punct } k = End;
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
  Lexer(const std::string& src, const LexOptions& opt) : s_(src), opt_(opt) {}
  std::vector<Token> run() {
    std::vector<Token> raw;
    for (;;) {
      Token t = next();
      raw.push_back(t);
      if (t.k == Token::End) break;
    }
    if (!opt_.glsl || macros_.empty()) return raw;
    std::vector<Token> out;
    expand(raw, out, 0);
    return out;
  }

 private:
  const std::string& s_;
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
      if (opt_.glsl && peek() == '#' && at_line_start_) { directive(); continue; }
      break;
    }
  }
  // handles #define (object- and function-like); ignores #version/#extension/#pragma/#line/precision
  void directive() {
    int l = line_, c = col_;
    std::string text;
    while (i_ < s_.size() && peek() != '\n') {
      if (peek() == '\\' && peek(1) == '\n') { adv(); adv(); text += ' '; continue; }
      text += peek();
      adv();
    }
    size_t p = 1;
    while (p < text.size() && isspace((unsigned char)text[p])) ++p;
    size_t q = p;
    while (q < text.size() && isalpha((unsigned char)text[q])) ++q;
    const std::string name = text.substr(p, q - p);
    if (name == "define") {
      while (q < text.size() && isspace((unsigned char)text[q])) ++q;
      size_t r = q;
      while (r < text.size() && (isalnum((unsigned char)text[r]) || text[r] == '_')) ++r;
      const std::string mname = text.substr(q, r - q);
      if (mname.empty()) err("#define without a name", l, c);
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
      const std::string body = text.substr(r);
      LexOptions o; o.glsl = false;
      Lexer sub(body, o);
      for (;;) { Token t = sub.next(); if (t.k == Token::End) break; t.line = l; t.col = c; m.body.push_back(t); }
      macros_[mname] = m;
    } else if (name == "if" || name == "ifdef" || name == "ifndef" || name == "else" || name == "elif" || name == "endif" ||
               name == "undef" || name == "include") {
      err("GLSL preprocessor directive #" + name + " is not supported", l, c);
    }
    // #version, #extension, #pragma, #line: ignored
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
