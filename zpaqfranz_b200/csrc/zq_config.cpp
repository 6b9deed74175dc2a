// zq_config.cpp -- method expansion, config generation, ZPAQL assembler (host, C++).
// See zq_config.h for the reference lines each part has to agree with. The ZPAQL programs below
// (post-processors and context hashers) are part of the ZPAQ *format contract*: their bytecode is
// stored in every block header, so the instruction sequences are fixed by bit-exactness.
#include "zq_config.h"

#include <cctype>
#include <cstdlib>
#include <cstring>

namespace zq {
namespace {

std::string num(long long x) { return std::to_string(x); }

int popcount_u(unsigned x) { return __builtin_popcount(x); }

// ---------------------------------------------------------------------------------------------
// ZPAQL token -> opcode. The numbering is the ZPAQ level-2 specification's:
//   L*8+{0:<>a 1:++ 2:-- 3:! 4:=0 7:=r}   L in a,b,c,d,*b,*c,*d   (39 jt, 47 jf, 55 r=a)
//   56 halt 57 out 59 hash 60 hashd 63 jmp
//   64+8*dst+src   dst=src                src 7 = immediate
//   128+8*op+src   a op= src              op in += -= *= /= %= &= &~ |= ^= <<= >>= == < >
//   255 lj
// Pseudo-ops (if/else/do/...) get codes >= 256.
enum Pseudo {
  P_POST = 256, P_PCOMP, P_END, P_IF, P_IFNOT, P_ELSE, P_ENDIF, P_DO, P_WHILE, P_UNTIL,
  P_FOREVER, P_IFL, P_IFNOTL, P_ELSEL, P_SEMI
};
enum { OP_JT = 39, OP_JF = 47, OP_JMP = 63, OP_LJ = 255 };

// parse "a","b","c","d","*b","*c","*d" at s; returns index 0..6 and advances, or -1
int parse_place(const char*& s) {
  const char* p = s;
  bool ind = false;
  if (*p == '*') ind = true, ++p;
  int r = -1;
  switch (*p) { case 'a': r = 0; break; case 'b': r = 1; break; case 'c': r = 2; break; case 'd': r = 3; break; }
  if (r < 0) return -1;
  if (ind) { if (r == 0) return -1; r += 3; }
  s = p + 1;
  return r;
}

int opcode_of(const std::string& tok_in) {
  std::string tok;
  for (char ch : tok_in) tok.push_back((char)tolower((unsigned char)ch));
  struct Named { const char* name; int code; };
  static const Named named[] = {
      {"error", 0}, {"halt", 56}, {"out", 57}, {"hash", 59}, {"hashd", 60}, {"jt", OP_JT},
      {"jf", OP_JF}, {"jmp", OP_JMP}, {"lj", OP_LJ}, {"r=a", 55}, {"post", P_POST},
      {"pcomp", P_PCOMP}, {"end", P_END}, {"if", P_IF}, {"ifnot", P_IFNOT}, {"else", P_ELSE},
      {"endif", P_ENDIF}, {"do", P_DO}, {"while", P_WHILE}, {"until", P_UNTIL},
      {"forever", P_FOREVER}, {"ifl", P_IFL}, {"ifnotl", P_IFNOTL}, {"elsel", P_ELSEL}, {";", P_SEMI}};
  for (const Named& nm : named) if (tok == nm.name) return nm.code;
  const char* s = tok.c_str();
  int L = parse_place(s);
  if (L < 0) return -1;
  std::string rest = s;
  if (rest == "++") return L * 8 + 1;
  if (rest == "--") return L * 8 + 2;
  if (rest == "!") return L * 8 + 3;
  if (rest == "=0") return L * 8 + 4;
  if (rest == "<>a") return L ? L * 8 : -1;
  if (rest == "=r") return L < 4 ? L * 8 + 7 : -1;
  if (L == 0) {  // a op= src
    static const char* ops[] = {"+=", "-=", "*=", "/=", "%=", "&=", "&~", "|=", "^=", "<<=", ">>=", "==", "<", ">"};
    int best = -1; size_t bestlen = 0;
    for (int k = 0; k < 14; ++k) {
      size_t len = strlen(ops[k]);
      if (rest.compare(0, len, ops[k]) == 0 && len > bestlen) best = k, bestlen = len;
    }
    if (best >= 0) {
      const char* q = rest.c_str() + bestlen;
      if (!*q) return 128 + best * 8 + 7;
      int src = parse_place(q);
      if (src >= 0 && !*q) return 128 + best * 8 + src;
      // fall through: could still be a plain assignment such as "a=b"
    }
  }
  if (!rest.empty() && rest[0] == '=') {
    const char* q = rest.c_str() + 1;
    if (!*q) return 64 + L * 8 + 7;
    int src = parse_place(q);
    if (src >= 0 && !*q) return 64 + L * 8 + src;
  }
  return -1;
}

struct Tokenizer {
  const char* p;
  int line = 1;
  int depth = 0;  // comment nesting
  explicit Tokenizer(const char* s) : p(s) {}
  // next token (whitespace separated, "(nested (comments))" skipped)
  std::string next() {
    bool in_tok_after_comment = false;
    for (; *p; ++p) {
      if (*p == '\n') ++line;
      if (*p == '(') { depth += 1; in_tok_after_comment = false; continue; }
      if (depth > 0) { if (*p == ')') --depth; continue; }
      if ((unsigned char)*p > ' ') break;
    }
    (void)in_tok_after_comment;
    if (!*p) throw Error("unexpected end of config");
    std::string t;
    while ((unsigned char)*p > ' ' && *p != '(') t.push_back(*p++);
    return t;
  }
  [[noreturn]] void fail(const std::string& what, const std::string& tok) {
    throw Error("Config line " + num(line) + " at " + tok + ": " + what);
  }
};

int number_token(Tokenizer& tz, const int* args, int lo, int hi) {
  std::string t = tz.next();
  int r = 0;
  if (t.size() >= 2 && t[0] == '$' && t[1] >= '1' && t[1] <= '9') {
    if (t.size() > 2 && t[2] == '+') r = atoi(t.c_str() + 3);
    if (args) r += args[t[1] - '1'];
  } else if (!t.empty() && (t[0] == '-' || isdigit((unsigned char)t[0]))) {
    r = atoi(t.c_str());
  } else {
    tz.fail("expected a number", t);
  }
  if (r < lo) tz.fail("number too low", t);
  if (r > hi) tz.fail("number too high", t);
  return r;
}

void expect_word(Tokenizer& tz, const char* w) {
  std::string t = tz.next();
  std::string l;
  for (char ch : t) l.push_back((char)tolower((unsigned char)ch));
  if (l != w) tz.fail(std::string("expected ") + w, t);
}

// Assemble one HCOMP/PCOMP body into `code` (trailing 0 appended). Returns the terminating pseudo-op.
int assemble_body(Tokenizer& tz, const int* args, std::vector<uint8_t>& code, int fixed_bytes) {
  std::vector<int> ifs, dos;
  auto pop = [&](std::vector<int>& st, const std::string& t) {
    if (st.empty()) tz.fail("unbalanced control structure", t);
    int v = st.back(); st.pop_back(); return v;
  };
  int op = 0;
  for (;;) {
    std::string t = tz.next();
    op = opcode_of(t);
    if (op < 0 || op == P_SEMI) tz.fail("unexpected", t);
    if (op == P_POST || op == P_PCOMP || op == P_END) break;
    int o1 = -1, o2 = -1;
    const int here = (int)code.size();
    switch (op) {
      case P_IF: op = OP_JF; o1 = 0; ifs.push_back(here + 1); break;
      case P_IFNOT: op = OP_JT; o1 = 0; ifs.push_back(here + 1); break;
      case P_IFL: case P_IFNOTL:
        code.push_back(op == P_IFL ? OP_JT : OP_JF); code.push_back(3);
        op = OP_LJ; o1 = o2 = 0; ifs.push_back((int)code.size() + 1);
        break;
      case P_ELSE: case P_ELSEL: {
        const bool lng = op == P_ELSEL;
        op = lng ? OP_LJ : OP_JMP; o1 = 0; if (lng) o2 = 0;
        int a = pop(ifs, t);  // operand slot of the matching if
        if (code[a - 1] != OP_LJ) {
          int j = (here - a) + 1 + (lng ? 1 : 0);
          if (j > 127) tz.fail("IF too big, try IFL, IFNOTL", t);
          code[a] = (uint8_t)j;
        } else {
          int j = here + 2 + (lng ? 1 : 0);
          code[a] = j & 255; code[a + 1] = (j >> 8) & 255;
        }
        ifs.push_back(here + 1);
        break;
      }
      case P_ENDIF: {
        int a = pop(ifs, t);
        if (code[a - 1] != OP_LJ) {
          int j = here - a - 1;
          if (j > 127) tz.fail("IF too big, try IFL, IFNOTL, ELSEL", t);
          code[a] = (uint8_t)j;
        } else {
          code[a] = here & 255; code[a + 1] = (here >> 8) & 255;
        }
        op = -1;
        break;
      }
      case P_DO: dos.push_back(here); op = -1; break;
      case P_WHILE: case P_UNTIL: case P_FOREVER: {
        int a = pop(dos, t);
        int j = a - here - 2;
        if (j >= -127) {
          op = op == P_WHILE ? OP_JT : op == P_UNTIL ? OP_JF : OP_JMP;
          o1 = j & 255;
        } else {
          if (op == P_WHILE) { code.push_back(OP_JF); code.push_back(3); }
          if (op == P_UNTIL) { code.push_back(OP_JT); code.push_back(3); }
          op = OP_LJ; o1 = a & 255; o2 = a >> 8;
        }
        break;
      }
      default:
        if ((op & 7) == 7) {
          if (op == OP_LJ) { int v = number_token(tz, args, 0, 65535); o1 = v & 255; o2 = v >> 8; }
          else if (op == OP_JT || op == OP_JF || op == OP_JMP) o1 = number_token(tz, args, -128, 127) & 255;
          else o1 = number_token(tz, args, 0, 255);
        }
    }
    if (op >= 0 && op <= 255) code.push_back((uint8_t)op);
    if (o1 >= 0) code.push_back((uint8_t)o1);
    if (o2 >= 0) code.push_back((uint8_t)o2);
    if ((int)code.size() + fixed_bytes + 128 >= 68000 - 130 || (int)code.size() + fixed_bytes - 2 > 65535)
      tz.fail("program too big", t);
  }
  code.push_back(0);
  return op;
}

const int kCompSize[10] = {0, 2, 3, 2, 3, 4, 6, 6, 3, 5};  // bytes per component descriptor (ZPAQ spec)
const char* kCompName[10] = {"", "const", "cm", "icm", "match", "avg", "mix2", "mix", "isse", "sse"};

// ---------------------------------------------------------------------------------------------
// Post-processor programs (inverse transforms, executed by the *decoder*; we only store them).
const char* kE8Body =
    "a=b a==d ifnot a+= 4 a<d if a=*b a&= 254 a== 232 if c=b b++ b++ b++ b++ a=*b a++ a&= 254 "
    "a== 0 if b-- a=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b a-=b a++ *b=a a>>= 8 b++ *b=a a>>= 8 b++ "
    "*b=a b++ endif b=c endif endif a=*b out b++ forever endif\n";

std::string pcomp_varlen(bool e8, int rb) {  // level 1 decoder
  std::string s = "pcomp lazy2 3 ;\n a> 255 if\n";
  if (e8) s += std::string(" b=0 d=r 4 do ") + kE8Body;
  s += " a=0 b=0 c=0 d=0 r=a 1 r=a 2 r=a 3 r=a 4 halt endif\n"
       " a<<=d a+=c c=a a= 8 a+=d d=a\n"
       " a=r 1 a== 0 if a= 1 r=a 2 a=c a&= 3 a> 0 if a-- a<<= 3 r=a 3 a=c a>>= 2 c=a b=r 3 a&= 7 "
       "a+=b r=a 3 a=c a>>= 3 c=a a=d a-= 5 d=a a= 1 r=a 1 else a=c a>>= 2 c=a d-- d-- a= 3 r=a 1 "
       "endif endif\n"
       " do a=r 1 a== 1 if a=d a> 2 if a=c a&= 1 a== 1 if a=c a>>= 1 c=a b=r 2 a=c a&= 1 a+=b a+=b "
       "r=a 2 a=c a>>= 1 c=a d-- d-- else a=c a>>= 1 c=a a=r 2 a<<= 2 b=a a=c a&= 3 a+=b r=a 2 a=c "
       "a>>= 2 c=a d-- d-- d-- ";
  s += rb ? "a= 5 r=a 1" : "a= 2 r=a 1";
  s += " endif forever endif endif\n";
  if (rb)
    s += " a=r 1 a== 5 if a=d a> " + num(rb - 1) + " if a=c a&= " + num((1 << rb) - 1) +
         " r=a 5 a=c a>>= " + num(rb) + " c=a a=d a-= " + num(rb) + " d=a a= 2 r=a 1 endif endif\n";
  s += " a=r 1 a== 2 if a=r 3 a>d ifnot a=c r=a 6 a=d r=a 7 b=r 3 a= 1 a<<=b d=a a-- a&=c a+=d\n";
  if (rb) s += " a<<= " + num(rb) + " d=r 5 a+=d a-= " + num((1 << rb) - 1) + "\n";
  s += " d=a b=r 4 a=b a-=d c=a d=r 2 do a=d a> 0 if d-- a=*c *b=a c++ b++";
  if (!e8) s += " out";
  s += " forever endif a=b r=a 4 a=r 6 b=r 3 a>>=b c=a a=r 7 a-=b d=a a=0 r=a 1 endif endif\n"
       " do a=r 1 a== 3 if a=d a> 1 if a=c a&= 1 a== 1 if a=c a>>= 1 c=a b=r 2 a&= 1 a+=b a+=b r=a 2 "
       "a=c a>>= 1 c=a d-- d-- else a=c a>>= 1 c=a d-- a= 4 r=a 1 endif forever endif endif\n"
       " a=r 1 a== 4 if a=d a> 7 if b=r 4 a=c *b=a";
  if (!e8) s += " out";
  s += " b++ a=b r=a 4 a=c a>>= 8 c=a a=d a-= 8 d=a a=r 2 a-- r=a 2 a== 0 if a=0 r=a 1 endif "
       "endif endif halt end\n";
  return s;
}

std::string pcomp_bytelz(bool e8) {  // level 2 decoder
  std::string s = "pcomp lzpre c ;\n a> 255 if\n";
  if (e8) s += std::string(" d=b b=0 do ") + kE8Body;
  s += " b=0 c=0 d=0 a=0 r=a 1 r=a 2 halt endif\n"
       " c=a a=d a== 0 if a=c a>>= 6 a++ d=a a== 1 if a+=c r=a 1 a=0 r=a 2 else d++ a=c a&= 63 "
       "a+= $3 r=a 1 a=0 r=a 2 endif else a== 1 if a=c *b=a b++";
  if (!e8) s += " out";
  s += " a=r 1 a-- a== 0 if d=0 endif r=a 1 else a> 2 if a=r 2 a<<= 8 a|=c r=a 2 d-- else a=r 2 "
       "a<<= 8 a|=c c=a a=b a-=c a-- c=a d=r 1 do a=*c *b=a c++ b++";
  if (!e8) s += " out";
  s += " d-- a=d a> 0 while endif endif endif halt end\n";
  return s;
}

std::string pcomp_ibwt(bool e8, int arg0) {  // level 3 decoder
  std::string s =
      "pcomp bwtrle c ;\n a> 255 ifnot *b=a b++ elsel\n"
      " b-- a=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b b-- a<<= 8 a+=*b c=a r=a 1 a=b r=a 2\n"
      " do a=b a> 0 if b-- a=*b a++ a&= 255 d=a d! *d++ forever endif\n"
      " d=0 d! *d= 1 a=0 do a+=*d *d=a d-- d<>a a! a> 255 a! d<>a until\n"
      " b=0 do a=c a>b if d=*b d! *d++ d=*d d-- *d=b b++ forever endif\n"
      " b=c b++ c=r 2 do a=c a>b if d=*b d! *d++ d=*d d-- *d=b b++ forever endif\n";
  if (arg0 <= 4) {
    s += " b=0 do a=c a>b if d=b a=*d a<<= 8 a+=*b *d=a b++ forever endif\n"
         " d=r 1 b=0 do a=d a== 0 ifnot a=*d a>>= 8 d=a";
    s += e8 ? " *b=*d b++" : " a=*d out";
    s += " forever endif\n";
    if (e8) s += std::string(" d=b b=0 do ") + kE8Body;
    s += " endif halt end\n";
  } else if (e8) {
    s += " a=r 2 a-- r=a 2 c=0 d=r 1 do a=d a== 0 ifnot d=*d b=d a=*b a<<= 24 b=a a=r 4 r=a 5 "
         "a>>= 8 a|=b r=a 4 a=c a> 3 if a=r 5 a&= 254 a== 232 if a=r 4 a>>= 24 b=a a++ a&= 254 "
         "a< 2 if a=r 4 a-=c a+= 4 a<<= 8 a>>= 8 b<>a a<<= 24 a+=b r=a 4 endif endif endif a=c "
         "a> 3 if a=r 5 out endif c++ forever endif\n"
         " b=r 4 a=c a> 3 a=b if out endif a>>= 8 b=a a=c a> 2 a=b if out endif a>>= 8 b=a a=c "
         "a> 1 a=b if out endif a>>= 8 b=a a=c a> 0 a=b if out endif endif halt end\n";
  } else {
    s += " d=r 1 do a=d a== 0 ifnot d=*d b=d a=*b out forever endif endif halt end\n";
  }
  return s;
}

const char* kPcompE8Only =
    "pcomp e8e9 d ;\n a> 255 if a=c a> 4 if c= 4 else a! a+= 5 a<<= 3 d=a a=b a>>=d b=a endif "
    "do a=c a> 0 if a=b out a>>= 8 b=a c-- forever endif else *b=b a<<= 24 d=a a=b a>>= 8 a+=d b=a "
    "c++ a=c a> 4 if a=*b out a&= 254 a== 232 if a=b a>>= 24 a++ a&= 254 a== 0 if a=b a>>= 24 "
    "a<<= 24 d=a a=b a-=c a+= 5 a<<= 8 a>>= 8 a|=d b=a endif endif endif endif halt end\n";

// One parsed model command: letter + numeric arguments (v[0] = the letter, v[1..] = its numbers: the format's own
// 1-based $N numbering).
typedef std::vector<int> Cmd;

// "x4,1,4,0,7,21,1c0,0,511i2..." -> args[0..8] and the list of model commands behind them
std::vector<Cmd> split_method(const std::string& method_s, int args[9]) {
  const char* m = method_s.c_str() + 1;
  for (int i = 0; i < 9; ++i) args[i] = 0;
  for (int i = 0; i < 9 && (isdigit((unsigned char)*m) || *m == ',' || *m == '.'); ++m) {
    if (isdigit((unsigned char)*m)) args[i] = args[i] * 10 + (*m - '0');
    else if (++i < 9) args[i] = 0;
  }
  std::vector<Cmd> cmds;
  while (*m) {
    Cmd v(1, (unsigned char)*m++);
    if (isdigit((unsigned char)*m)) {
      v.push_back(*m++ - '0');
      while (isdigit((unsigned char)*m) || *m == ',' || *m == '.') {
        if (isdigit((unsigned char)*m)) v.back() = v.back() * 10 + (*m++ - '0');
        else { v.push_back(0); ++m; }
      }
    }
    cmds.push_back(v);
  }
  return cmds;
}

// Grows the COMP list and the HCOMP program of a model, one method command at a time.  Conventions of every generated
// HCOMP (they are part of the stored header, hence of the format): M = backwards-filling history of the last 64 KiB,
// C -> newest byte, H[0..254] = component contexts, H[255+b] = position of byte b's last occurrence; for byte-LZ77 R1/R2
// track the parse state of the code stream.
struct ModelBuilder {
  int membits, ncomp = 0, sb = 5;
  std::string comp, hc;

  void context(Cmd v) {   // 'c': ICM or CM over a hashed (periodic | distance | masked-byte) context
    while (v.size() < 3) v.push_back(0);
    sb = 11 + (v[2] < 256 ? bitlen((unsigned)v[2]) : 6);
    for (size_t i = 3; i < v.size(); ++i) if (v[i] < 512) sb += popcount_u((unsigned)v[i]) * 3 / 4;
    if (sb > membits) sb = membits;
    comp += num(ncomp) + " ";
    if (v[1] % 1000 == 0) comp += "icm " + num(sb - 6 - v[1] / 1000) + "\n";
    else comp += "cm " + num(sb - 2 - v[1] / 1000) + " " + num(v[1] % 1000 - 1) + "\n";
    hc += "d= " + num(ncomp) + " *d=0\n";
    if (v[2] > 1 && v[2] <= 255) {
      const bool pow2 = bitlen((unsigned)v[2]) != bitlen((unsigned)v[2] - 1);
      hc += pow2 ? "a=c a&= " + num(v[2] - 1) + " hashd\n" : "a=c a%= " + num(v[2]) + " hashd\n";
    } else if (v[2] >= 1000 && v[2] <= 1255) {
      hc += "a= 255 a+= " + num(v[2] - 1000) + " d=a a=*d a-=c a> 255 if a= 255 endif d= " + num(ncomp) + " hashd\n";
    }
    for (size_t i = 3; i < v.size(); ++i) {
      const int x = v[i];
      if (i == 3) hc += "b=c ";
      if (x == 255) hc += "a=*b hashd\n";
      else if (x > 0 && x < 255) hc += "a=*b a&= " + num(x) + " hashd\n";
      else if (x >= 256 && x < 512) {
        hc += "a=r 1 a> 1 if a=r 2 a< 64 if a=*b ";
        if (x < 511) hc += "a&= " + num(x - 256);
        hc += " hashd else a>>= 6 hashd a=r 1 hashd endif else a= 255 hashd a=r 2 hashd endif\n";
      }
      else if (x >= 1256) hc += "a= " + num(((x - 1000) >> 8) & 255) + " a<<= 8 a+= " + num((x - 1000) & 255) + " a+=b b=a\n";
      else if (x > 1000) hc += "a= " + num(x - 1000) + " a+=b b=a\n";
      if (i + 1 < v.size() && x < 512) hc += "b++ ";
    }
    ++ncomp;
  }

  void combiner(Cmd v) {   // 'm' MIX, 't' MIX2, 's' SSE over the components so far
    const int L = v[0];
    if (ncomp <= (L == 't' ? 1 : 0)) return;
    if (v.size() <= 1) v.push_back(8);
    if (v.size() <= 2) v.push_back(24 + 8 * (L == 's'));
    if (L == 's' && v.size() <= 3) v.push_back(255);
    comp += num(ncomp);
    sb = 5 + v[1] * 3 / 4;
    if (L == 'm') comp += " mix " + num(v[1]) + " 0 " + num(ncomp) + " " + num(v[2]) + " 255\n";
    else if (L == 't') comp += " mix2 " + num(v[1]) + " " + num(ncomp - 1) + " " + num(ncomp - 2) + " " + num(v[2]) + " 255\n";
    else comp += " sse " + num(v[1]) + " " + num(ncomp - 1) + " " + num(v[2]) + " " + num(v[3]) + "\n";
    if (v[1] > 8) {  // order-1/2 byte context shifted past the 8 partial-byte bits
      hc += "d= " + num(ncomp) + " *d=0 b=c a=0\n";
      int bits = v[1];
      for (; bits >= 16; bits -= 8) { hc += "a<<= 8 a+=*b"; if (bits > 16) hc += " b++"; hc += "\n"; }
      if (bits > 8) hc += "a<<= 8 a+=*b a>>= " + num(16 - bits) + "\n";
      hc += "a<<= 8 *d=a\n";
    }
    ++ncomp;
  }

  void isse_chain(Cmd v) {   // 'i': each link extends the previous hash by N bytes
    if (ncomp <= 0) return;
    hc += "d= " + num(ncomp - 1) + " b=c a=*d d++\n";
    for (size_t i = 1; i < v.size() && ncomp < 254; ++i) {
      for (int j = 0; j < v[i] % 10; ++j) {
        hc += "hash ";
        if (i + 1 < v.size() || j < v[i] % 10 - 1) hc += "b++ ";
        sb += 6;
      }
      hc += "*d=a";
      if (i + 1 < v.size()) hc += " d++";
      hc += "\n";
      if (sb > membits) sb = membits;
      comp += num(ncomp) + " isse " + num(sb - 6 - v[i] / 10) + " " + num(ncomp - 1) + "\n";
      ++ncomp;
    }
  }

  void match(Cmd v) {   // 'a'
    if (v.size() <= 1) v.push_back(24);
    while (v.size() < 4) v.push_back(0);
    comp += num(ncomp) + " match " + num(membits - v[3] - 2) + " " + num(membits - v[2]) + "\n";
    hc += "d= " + num(ncomp) + " a=*d a*= " + num(v[1]) + " a+=*c a++ *d=a\n";
    sb = 5 + (membits - v[2]) * 3 / 4;
    ++ncomp;
  }

  void words(Cmd v) {   // 'w': word-oriented ICM-ISSE chain
    static const int dflt[7] = {0, 1, 65, 26, 223, 20, 0};
    for (size_t k = v.size(); k <= 6; ++k) v.push_back(dflt[k]);
    comp += num(ncomp) + " icm " + num(membits - 6 - v[6]) + "\n";
    for (int i = 1; i < v[1]; ++i)
      comp += num(ncomp + i) + " isse " + num(membits - 6 - v[6]) + " " + num(ncomp + i - 1) + "\n";
    hc += "a=*c a&= " + num(v[4]) + " a-= " + num(v[2]) + " a&= 255 a< " + num(v[3]) + " if\n";
    for (int i = 0; i < v[1]; ++i)
      hc += (i == 0 ? "  d= " + num(ncomp) : std::string("  d++")) + " a=*d a*= " + num(v[5]) + " a+=*c a++ *d=a\n";
    hc += "else\n";
    for (int i = v[1] - 1; i > 0; --i) hc += "  d= " + num(ncomp + i - 1) + " a=*d d++ *d=a\n";
    hc += "  d= " + num(ncomp) + " *d=0\nendif\n";
    ncomp += v[1] - 1;
    sb = membits - v[6];
    ++ncomp;
  }
};

// method letter -> what it adds to the model
const struct { char letter; void (ModelBuilder::*add)(Cmd); } kModelCommands[] = {
    {'c', &ModelBuilder::context}, {'m', &ModelBuilder::combiner}, {'t', &ModelBuilder::combiner}, {'s', &ModelBuilder::combiner},
    {'i', &ModelBuilder::isse_chain}, {'a', &ModelBuilder::match}, {'w', &ModelBuilder::words},
};

}  // namespace

// -------------------------------------------------------------------------------------------------
// Method string -> ZPAQL source of the block's model and post-processor (the job of makeConfig, Z:19615): a pre-pass
// header chosen by args[1], then one table-driven step per model command.
std::string make_config(const std::string& method_s, int args[9]) {
  const char type = method_s.c_str()[0];
  if (!(type == 'x' || type == 's' || type == '0' || type == 'i')) throw Error("Unsupported method");
  const std::vector<Cmd> cmds = split_method(method_s, args);
  if (type == '0') return "comp 0 0 0 0 0 hcomp end\n";

  const int lz = args[1] & 3;
  const bool e8 = args[1] >= 4 && args[1] <= 7;
  std::string hdr, post;
  if (lz == 1) { hdr = "comp 9 16 0 $1+20 "; post = pcomp_varlen(e8, args[0] > 4 ? args[0] - 4 : 0); }
  else if (lz == 2) { hdr = "comp 9 16 0 $1+20 "; post = pcomp_bytelz(e8); }
  else if (lz == 3) { hdr = "comp 9 16 $1+20 $1+20 "; post = pcomp_ibwt(e8, args[0]); }
  else { hdr = "comp 9 16 0 0 "; post = e8 ? kPcompE8Only : "end\n"; }

  ModelBuilder mb;
  mb.membits = args[0] + 20;
  mb.hc = "hcomp\nc-- *c=a a+= 255 d=a *d=c\n";
  if (lz == 2)
    mb.hc += "a=r 1 a== 0 if a= " + num(111 + 57 * (e8 ? 1 : 0)) +
             " else a== 1 if a=*c r=a 2 a> 63 if a>>= 6 a++ a++ else a++ a++ endif else a-- endif endif r=a 1\n";
  for (const Cmd& c : cmds) {
    if (mb.ncomp >= 254) break;
    for (const auto& k : kModelCommands)
      if (k.letter == (char)c[0]) { (mb.*k.add)(c); break; }
  }
  return hdr + num(mb.ncomp) + "\n" + mb.comp + mb.hc + "halt\n" + post;
}

// -------------------------------------------------------------------------------------------------
Assembled assemble(const std::string& config, const int args[9]) {
  Assembled r;
  Tokenizer tz(config.c_str());
  expect_word(tz, "comp");
  r.hh = number_token(tz, args, 0, 255);
  r.hm = number_token(tz, args, 0, 255);
  r.ph = number_token(tz, args, 0, 255);
  r.pm = number_token(tz, args, 0, 255);
  r.ncomp = number_token(tz, args, 0, 255);
  for (int i = 0; i < r.ncomp; ++i) {
    number_token(tz, args, i, i);
    std::string t = tz.next(), l;
    for (char ch : t) l.push_back((char)tolower((unsigned char)ch));
    int type = -1;
    for (int k = 1; k < 10; ++k) if (l == kCompName[k]) type = k;
    if (type < 0) tz.fail("unexpected", t);
    r.comp.push_back((uint8_t)type);
    for (int j = 1; j < kCompSize[type]; ++j) r.comp.push_back((uint8_t)number_token(tz, args, 0, 255));
  }
  const int cend = 7 + (int)r.comp.size() + 1;
  expect_word(tz, "hcomp");
  int endop = assemble_body(tz, args, r.hcomp, cend);
  const int hsize = (cend - 2) + (int)r.hcomp.size();
  r.header.push_back(hsize & 255); r.header.push_back(hsize >> 8);
  r.header.push_back(r.hh); r.header.push_back(r.hm); r.header.push_back(r.ph); r.header.push_back(r.pm);
  r.header.push_back(r.ncomp);
  r.header.insert(r.header.end(), r.comp.begin(), r.comp.end());
  r.header.push_back(0);
  r.header.insert(r.header.end(), r.hcomp.begin(), r.hcomp.end());
  if (endop == P_POST) {
    number_token(tz, args, 0, 0);
    expect_word(tz, "end");
  } else if (endop == P_PCOMP) {
    // command text up to ';' (kept verbatim, case sensitive)
    while (*tz.p && (unsigned char)*tz.p <= ' ') { if (*tz.p == '\n') ++tz.line; ++tz.p; }
    while (*tz.p && *tz.p != ';') r.pcomp_cmd.push_back(*tz.p++);
    if (*tz.p) ++tz.p;
    endop = assemble_body(tz, args, r.pcomp, 8);
    if (endop != P_END) throw Error("Config: expected END");
  } else if (endop != P_END) {
    throw Error("Config: expected END or POST 0 END or PCOMP cmd ; ... END");
  }
  return r;
}

// -------------------------------------------------------------------------------------------------
std::string expand_method(const std::string& method_in, const uint8_t* data, uint32_t n) {
  if (method_in.empty()) throw Error("empty method");
  if (!isdigit((unsigned char)method_in[0])) return method_in;
  int arg0 = bitlen(n + 4095) - 20; if (arg0 < 0) arg0 = 0;
  // "LB,R,t": R = redundancy 0..255, t = 1 text | 2 exe. No suffix => type 512.
  unsigned type = 512;
  {
    int commas = 0, f[4] = {0, 0, 0, 0};
    for (size_t i = 1; i < method_in.size() && commas < 4; ++i) {
      const char ch = method_in[i];
      if (ch == ',' || ch == '.') ++commas;
      else if (isdigit((unsigned char)ch)) f[commas] = f[commas] * 10 + (ch - '0');
    }
    if (commas) type = (unsigned)(f[1] * 4 + f[2]);
  }
  const int level = method_in[0] - '0';
  const int e8 = (type & 2) * 2;  // +4 on the pre-pass selector when the block looks like x86 code
  const std::string X = "x" + num(arg0);
  const std::string ht = "," + num(19 + arg0 + (arg0 <= 6 ? 1 : 0));  // LZ77 hash table bits
  const std::string sa = "," + num(21 + arg0);                         // => suffix-array search
  auto lz = [&](int kind) { return X + "," + num(kind + e8) + ","; };
  switch (level) {
    case 0: return "0" + num(arg0) + ",0";
    case 1:
      if (type < 40) return X + ",0";
      if (type < 80) return lz(1) + "4,0,1,15";
      if (type < 128) return lz(1) + "4,0,2,16";
      if (type < 256) return lz(1) + "4,0,2" + ht;
      if (type < 960) return lz(1) + "5,0,3" + ht;
      return lz(1) + "6,0,3" + ht;
    case 2:
      if (type < 32) return X + ",0";
      if (type < 64) return lz(1) + "4,0,3" + ht;
      return lz(1) + "4,0,7" + sa + ",1";
    case 3:
      if (type < 20) return X + ",0";
      if (type < 48) return lz(1) + "4,0,3" + ht;
      if (type >= 640 || (type & 1)) return X + "," + num(3 + e8) + "ci1";
      return lz(2) + "12,0,7" + sa + ",1c0,0,511i2";
    case 4:
      if (type < 12) return X + ",0";
      if (type < 24) return lz(1) + "4,0,3" + ht;
      if (type < 48) return lz(2) + "5,0,7" + sa + "1c0,0,511";  // (sic) upstream concatenation
      if (type < 900) return X + "," + num(e8) + "ci1,1,1,1,2a" + ((type & 1) ? "w" : "") + "m";
      return X + "," + num(3 + e8) + "ci1";
    default: break;
  }
  // levels 5..9: many-model CM; up to two periodic contexts from the byte-gap histogram
  int periods[2] = {0, 0};
  gap_periods(data, n, periods);
  return expand_method_periods(method_in, n, periods);
}

void gap_periods(const uint8_t* data, uint32_t n, int periods[2]) {
  periods[0] = periods[1] = 0;
  const int NR = 1 << 12;
  std::vector<int> gap(NR, 0);
  int last[256] = {0};
  for (uint32_t i = 0; i < n; ++i) {
    const int k = (int)i - last[data[i]];
    if (k > 0 && k < NR) ++gap[k];
    last[data[i]] = (int)i;
  }
  int n1 = (int)n - gap[1] - gap[2] - gap[3];
  for (int rep = 0; rep < 2; ++rep) {
    int period = 0, t = 0;
    double best = 0;
    for (int j = 5; j < NR && t < n1; ++j) {
      const double s = gap[j] / (256.0 + n1 - t);
      if (s > best) best = s, period = j;
      t += gap[j];
    }
    if (!(period > 4 && best > 0.1)) break;
    periods[rep] = period;
    n1 -= gap[period];
    gap[period] = 0;
  }
}

std::string expand_method_periods(const std::string& method_in, uint32_t n, const int periods[2]) {
  int arg0 = bitlen(n + 4095) - 20; if (arg0 < 0) arg0 = 0;
  unsigned type = 512;
  {
    int commas = 0, f[4] = {0, 0, 0, 0};
    for (size_t i = 1; i < method_in.size() && commas < 4; ++i) {
      const char ch = method_in[i];
      if (ch == ',' || ch == '.') ++commas;
      else if (isdigit((unsigned char)ch)) f[commas] = f[commas] * 10 + (ch - '0');
    }
    if (commas) type = (unsigned)(f[1] * 4 + f[2]);
  }
  const int e8 = (type & 2) * 2;
  std::string mth = "x" + num(arg0) + "," + num(e8) + ((type & 1) ? "w2c0,1010,255i1" : "w1i1") + "c256ci1,1,1,1,1,1,2a";
  for (int rep = 0; rep < 2 && periods[rep] > 4; ++rep) {
    mth += "c0,0," + num(999 + periods[rep]) + ",255i1";
    if (periods[rep] <= 255) mth += "c0," + num(periods[rep]) + "i1";
  }
  return mth + "c0,2,0,255i1c0,3,0,0,255i1c0,4,0,0,0,255i1mm16ts19t0";
}

BlockPlan plan_block(const std::string& method, const uint8_t* data, uint32_t n) {
  BlockPlan p;
  p.method = expand_method(method, data, n);
  std::string cfg = make_config(p.method, p.args);
  p.code = assemble(cfg, p.args);
  p.stored = p.method[0] == '0';
  const int a1 = p.args[1];
  const bool pre = a1 >= 1 && a1 <= 7 && a1 != 4;
  p.lz_level = pre ? (a1 & 3) : 0;
  p.e8e9 = a1 >= 4 && a1 <= 7;
  p.use_sa = pre && (p.lz_level == 3 || p.args[5] - p.args[0] >= 21);
  return p;
}

}  // namespace zq
