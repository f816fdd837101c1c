// wgsl_emit.cpp -- WGSL -> CUDA C++ emitter.
//
// Replaces naga-cranelift on the GPU path: where the reference JIT-compiles naga IR to x86
// (naga-cranelift/src/lib.rs:83-113, compiler.rs:192-608) this front end turns one WGSL entry
// point into CUDA C++ that is inlined into the raster kernels (wgb_raster.cuh) and compiled with
// NVRTC.  The Rust host would feed naga IR to an emitter of the same output format; Rust is not
// available in this image, so the front end parses WGSL itself.
//
// Coverage: everything the reference's JIT implements (literals, constants,
// zero values, compose, access, locals, globals, load/store, textureSample, unary, binary,
// select, casts, calls; block/if/switch/loop/break/continue/return/discard/store/call) plus
// what it leaves as todo!(): swizzles, splats, math builtins, vector comparisons, matrix*matrix.
// The arithmetic contract is the reference's: one IEEE binary32 operation per operator, never
// contracted (every float operator is emitted as a wgb_* call or a prelude operator built on the
// round-to-nearest intrinsics); mat*vec accumulates columns left to right (binary.rs:297-323).
// Deliberate deviations from the reference's quirks: `break if` has WGSL polarity (the reference
// inverts it, loop.rs:67-73), integer division by zero yields WGSL's defined result instead of an
// abort, float->int casts saturate instead of trapping.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

[[noreturn]] void err(int line, const std::string& m) {
    throw std::runtime_error("WGSL:" + std::to_string(line) + ": " + m);
}

// ------------------------------------------------------------------------------------------
// lexer
// ------------------------------------------------------------------------------------------
struct Tok {
    enum K { End, Ident, Int, Float, Punct, Attr } k = End;
    std::string s;      // identifier / punctuation / attribute name / literal text
    double f = 0;       // numeric value
    char suffix = 0;    // i u f h or 0
    int line = 1;
};

std::vector<Tok> lex(const std::string& src) {
    std::vector<Tok> out;
    size_t i = 0, n = src.size();
    int line = 1;
    auto isid0 = [](char c) { return isalpha((unsigned char)c) || c == '_'; };
    auto isid = [](char c) { return isalnum((unsigned char)c) || c == '_'; };
    while (i < n) {
        const char c = src[i];
        if (c == '\n') { line++; i++; continue; }
        if (isspace((unsigned char)c)) { i++; continue; }
        if (c == '/' && i + 1 < n && src[i + 1] == '/') { while (i < n && src[i] != '\n') i++; continue; }
        if (c == '/' && i + 1 < n && src[i + 1] == '*') {
            int depth = 1; i += 2;
            while (i < n && depth > 0) {
                if (src[i] == '\n') line++;
                if (src[i] == '/' && i + 1 < n && src[i + 1] == '*') { depth++; i += 2; }
                else if (src[i] == '*' && i + 1 < n && src[i + 1] == '/') { depth--; i += 2; }
                else i++;
            }
            continue;
        }
        Tok t; t.line = line;
        if (c == '@') {
            size_t j = i + 1;
            while (j < n && isspace((unsigned char)src[j])) j++;
            size_t k = j;
            while (k < n && isid(src[k])) k++;
            t.k = Tok::Attr; t.s = src.substr(j, k - j);
            i = k; out.push_back(t); continue;
        }
        if (isid0(c)) {
            size_t k = i;
            while (k < n && isid(src[k])) k++;
            t.k = Tok::Ident; t.s = src.substr(i, k - i);
            i = k; out.push_back(t); continue;
        }
        if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)src[i + 1]))) {
            size_t k = i;
            bool isf = false, hex = false;
            if (c == '0' && k + 1 < n && (src[k + 1] == 'x' || src[k + 1] == 'X')) {
                hex = true; k += 2;
                while (k < n && (isxdigit((unsigned char)src[k]) || src[k] == '.')) { if (src[k] == '.') isf = true; k++; }
                if (k < n && (src[k] == 'p' || src[k] == 'P')) { isf = true; k++; if (k < n && (src[k] == '+' || src[k] == '-')) k++; while (k < n && isdigit((unsigned char)src[k])) k++; }
            } else {
                while (k < n && isdigit((unsigned char)src[k])) k++;
                if (k < n && src[k] == '.') { isf = true; k++; while (k < n && isdigit((unsigned char)src[k])) k++; }
                if (k < n && (src[k] == 'e' || src[k] == 'E')) {
                    size_t m = k + 1;
                    if (m < n && (src[m] == '+' || src[m] == '-')) m++;
                    if (m < n && isdigit((unsigned char)src[m])) { isf = true; k = m; while (k < n && isdigit((unsigned char)src[k])) k++; }
                }
            }
            const std::string text = src.substr(i, k - i);
            if (k < n && (src[k] == 'f' || src[k] == 'h') && !hex) { t.suffix = src[k]; isf = true; k++; }
            else if (k < n && (src[k] == 'i' || src[k] == 'u')) { t.suffix = src[k]; k++; }
            t.k = isf ? Tok::Float : Tok::Int;
            t.s = text;
            t.f = isf ? strtod(text.c_str(), nullptr) : (double)strtoull(text.c_str(), nullptr, 0);
            i = k; out.push_back(t); continue;
        }
        static const char* puncts[] = {"<<=", ">>=", "->", "==", "!=", "<=", ">=", "&&", "||", "<<", ">>", "+=", "-=", "*=", "/=", "%=",
                                        "&=", "|=", "^=", "++", "--", nullptr};
        bool matched = false;
        for (int p = 0; puncts[p]; p++) {
            const size_t L = strlen(puncts[p]);
            if (src.compare(i, L, puncts[p]) == 0) { t.k = Tok::Punct; t.s = puncts[p]; i += L; matched = true; break; }
        }
        if (!matched) { t.k = Tok::Punct; t.s = std::string(1, c); i++; }
        out.push_back(t);
    }
    Tok e; e.k = Tok::End; e.line = line;
    out.push_back(e);
    return out;
}

// ------------------------------------------------------------------------------------------
// types
// ------------------------------------------------------------------------------------------
struct StructDecl;
struct Type {
    enum K { Void, Bool, I32, U32, F32, F16, AInt, AFloat, Vec, Mat, Array, Struct, Texture2D, Sampler } k = Void;
    int n = 0;            // vector size / matrix columns
    int rows = 0;         // matrix rows
    std::shared_ptr<Type> elem;   // vector / matrix scalar, array element
    int count = 0;        // array length (0 = runtime sized)
    const StructDecl* st = nullptr;

    bool is_scalar() const { return k == Bool || k == I32 || k == U32 || k == F32 || k == F16 || k == AInt || k == AFloat; }
    bool is_abstract() const { return k == AInt || k == AFloat || ((k == Vec || k == Mat || k == Array) && elem && elem->is_abstract()); }
    bool is_float_scalar() const { return k == F32 || k == F16 || k == AFloat; }
    bool is_int_scalar() const { return k == I32 || k == U32 || k == AInt; }
    const Type& scalar() const { return (k == Vec || k == Mat) ? *elem : *this; }
};
Type T(Type::K k) { Type t; t.k = k; return t; }
Type vec_t(int n, Type e) { Type t; t.k = Type::Vec; t.n = n; t.elem = std::make_shared<Type>(e); return t; }
Type mat_t(int c, int r, Type e) { Type t; t.k = Type::Mat; t.n = c; t.rows = r; t.elem = std::make_shared<Type>(e); return t; }
Type array_t(Type e, int count) { Type t; t.k = Type::Array; t.count = count; t.elem = std::make_shared<Type>(e); return t; }

bool same(const Type& a, const Type& b) {
    if (a.k != b.k) return false;
    switch (a.k) {
        case Type::Vec: return a.n == b.n && same(*a.elem, *b.elem);
        case Type::Mat: return a.n == b.n && a.rows == b.rows && same(*a.elem, *b.elem);
        case Type::Array: return a.count == b.count && same(*a.elem, *b.elem);
        case Type::Struct: return a.st == b.st;
        default: return true;
    }
}

struct Attr { std::string name; std::vector<std::string> args; };
struct Member { std::string name; Type type; std::vector<Attr> attrs; uint32_t offset = 0; int line = 0; };
struct StructDecl { std::string name; std::vector<Member> members; uint32_t size = 0, align = 0; };

std::string scalar_suffix(const Type& s, int line) {
    switch (s.k) {
        case Type::F32: case Type::AFloat: return "f";
        case Type::F16: return "h";
        case Type::I32: case Type::AInt: return "i";
        case Type::U32: return "u";
        case Type::Bool: return "b";
        default: err(line, "unsupported scalar type");
    }
}
std::string cuda_type(const Type& t, int line) {
    switch (t.k) {
        case Type::Void: return "void";
        case Type::Bool: return "bool";
        case Type::I32: case Type::AInt: return "i32";
        case Type::U32: return "u32";
        case Type::F32: case Type::AFloat: return "f32";
        case Type::F16: return "f16";
        case Type::Vec: return "vec" + std::to_string(t.n) + scalar_suffix(*t.elem, line);
        case Type::Mat:
            if (t.n != t.rows || !(t.n == 2 || t.n == 3 || t.n == 4)) err(line, "only square matrices are supported");
            if (t.elem->k == Type::F16) err(line, "f16 matrices are not supported");
            return "mat" + std::to_string(t.n) + "x" + std::to_string(t.rows) + "f";
        case Type::Array:
            if (t.count == 0) err(line, "runtime-sized arrays cannot be used by value");
            return "wgb_array<" + cuda_type(*t.elem, line) + ", " + std::to_string(t.count) + ">";
        case Type::Struct: return t.st->name;
        default: err(line, "resource types cannot be used by value");
    }
}
std::string wgsl_type_name(const Type& t) {
    switch (t.k) {
        case Type::Void: return "void"; case Type::Bool: return "bool"; case Type::I32: return "i32"; case Type::U32: return "u32";
        case Type::F32: return "f32"; case Type::F16: return "f16"; case Type::AInt: return "abstract-int"; case Type::AFloat: return "abstract-float";
        case Type::Vec: return "vec" + std::to_string(t.n) + "<" + wgsl_type_name(*t.elem) + ">";
        case Type::Mat: return "mat" + std::to_string(t.n) + "x" + std::to_string(t.rows) + "<" + wgsl_type_name(*t.elem) + ">";
        case Type::Array: return "array<" + wgsl_type_name(*t.elem) + ">";
        case Type::Struct: return t.st->name;
        case Type::Texture2D: return "texture_2d<f32>";
        default: return "sampler";
    }
}

// WGSL memory layout (alignment, size) of host-shareable types
void layout_of(const Type& t, uint32_t& align, uint32_t& size, int line) {
    switch (t.k) {
        case Type::I32: case Type::U32: case Type::F32: align = 4; size = 4; return;
        case Type::F16: err(line, "f16 in uniform / storage buffers is not supported (f16 values live in function and private variables)");
        case Type::Vec:
            if (t.elem->k == Type::F16) err(line, "f16 in uniform / storage buffers is not supported (f16 values live in function and private variables)");
            align = t.n == 2 ? 8 : 16; size = 4 * t.n; return;
        case Type::Mat: { const uint32_t ca = t.rows == 2 ? 8 : 16; align = ca; size = ca * t.n; return; }
        case Type::Array: {
            uint32_t ea, es;
            layout_of(*t.elem, ea, es, line);
            const uint32_t stride = (es + ea - 1) / ea * ea;
            align = ea; size = stride * (t.count ? t.count : 1);
            return;
        }
        case Type::Struct: align = t.st->align; size = t.st->size; return;
        default: err(line, "type " + wgsl_type_name(t) + " is not host-shareable");
    }
}

// ------------------------------------------------------------------------------------------
// AST
// ------------------------------------------------------------------------------------------
struct Expr;
using ExprP = std::shared_ptr<Expr>;
struct Expr {
    enum K { Lit, Ident, Unary, Binary, Call, Member, Index, Construct } k;
    int line = 0;
    // literal
    double f = 0; bool is_float = false; char suffix = 0; bool is_bool = false;
    std::string name;                 // ident / member / callee / operator
    std::vector<ExprP> args;          // operands
    Type ctor;                        // Construct: target type (elem may be absent: inferred, e.g. vec3(...))
    bool ctor_infer = false;
    Type type;                        // resolved by the checker
};
struct Stmt;
using StmtP = std::shared_ptr<Stmt>;
struct Stmt {
    enum K { Block, Let, Var, Const, Assign, Incr, If, For, While, Loop, Switch, Break, Continue, Return, Discard, CallS, BreakIf } k;
    int line = 0;
    std::string name, op;             // declared name / assignment operator
    bool has_type = false; Type decl_type;
    ExprP a, b;                       // init / lhs, rhs / condition / return value
    std::vector<StmtP> body, else_body, continuing;
    StmtP init, update;               // for
    struct Case { std::vector<ExprP> sel; bool is_default = false; std::vector<StmtP> body; };
    std::vector<Case> cases;
};
struct Param { std::string name; Type type; std::vector<Attr> attrs; int line = 0; };
struct Function {
    std::string name; std::vector<Param> params; Type ret; std::vector<Attr> ret_attrs; std::vector<Attr> attrs;
    std::vector<StmtP> body; int line = 0;
    bool may_discard = false;
    int stage() const { for (auto& a : attrs) { if (a.name == "vertex") return 1; if (a.name == "fragment") return 2; if (a.name == "compute") return 3; } return 0; }
};
struct Global {
    std::string name; Type type; std::string space;   // "uniform" "storage" "private" "handle" "const"
    int group = -1, binding = -1; ExprP init; int line = 0;
};
struct Module {
    std::vector<std::shared_ptr<StructDecl>> structs;
    std::map<std::string, Type> aliases;
    std::vector<Global> globals;
    std::vector<std::shared_ptr<Function>> functions;
    std::vector<std::string> decl_order;   // "s:<name>" "g:<name>" "f:<name>"
};

// ------------------------------------------------------------------------------------------
// parser
// ------------------------------------------------------------------------------------------
struct Parser {
    std::vector<Tok> t;
    size_t p = 0;
    Module m;

    const Tok& cur() const { return t[p]; }
    const Tok& peek(int d = 1) const { return t[std::min(p + d, t.size() - 1)]; }
    bool is(const char* s) const { return (cur().k == Tok::Punct || cur().k == Tok::Ident) && cur().s == s; }
    bool accept(const char* s) { if (is(s)) { p++; return true; } return false; }
    void expect(const char* s) { if (!accept(s)) err(cur().line, std::string("expected '") + s + "' but found '" + cur().s + "'"); }
    std::string ident() { if (cur().k != Tok::Ident) err(cur().line, "expected identifier, found '" + cur().s + "'"); return t[p++].s; }

    std::vector<Attr> attrs() {
        std::vector<Attr> out;
        while (cur().k == Tok::Attr) {
            Attr a; a.name = cur().s; p++;
            if (accept("(")) {
                int depth = 0; std::string argtxt;
                while (!(is(")") && depth == 0)) {
                    if (cur().k == Tok::End) err(cur().line, "unterminated attribute");
                    if (is("(")) depth++;
                    if (is(")")) depth--;
                    if (is(",") && depth == 0) { a.args.push_back(argtxt); argtxt.clear(); p++; continue; }
                    argtxt += cur().s; p++;
                }
                if (!argtxt.empty()) a.args.push_back(argtxt);
                expect(")");
            }
            out.push_back(a);
        }
        return out;
    }

    Type scalar_by_name(const std::string& s, int line) {
        if (s == "f32") return T(Type::F32);
        if (s == "i32") return T(Type::I32);
        if (s == "u32") return T(Type::U32);
        if (s == "bool") return T(Type::Bool);
        if (s == "f16") return T(Type::F16);
        err(line, "unknown scalar type '" + s + "'");
    }
    // parses a type; returns false in *ok if the identifier is not a type name
    bool try_type(Type& out, bool allow_infer = false, bool* inferred = nullptr) {
        if (cur().k != Tok::Ident) return false;
        const std::string s = cur().s;
        const int line = cur().line;
        auto templ_scalar = [&](Type& e) -> bool {
            if (accept("<")) { Type x; if (!try_type(x)) err(line, "expected type argument"); e = x; close_angle(); return true; }
            return false;
        };
        if (s == "f32" || s == "i32" || s == "u32" || s == "bool" || s == "f16") { p++; out = scalar_by_name(s, line); return true; }
        if (s.size() >= 4 && s.compare(0, 3, "vec") == 0 && isdigit((unsigned char)s[3])) {
            const int n = s[3] - '0';
            if (n < 2 || n > 4) return false;
            if (s.size() == 4) {
                p++;
                Type e;
                if (templ_scalar(e)) { out = vec_t(n, e); return true; }
                if (!allow_infer) err(line, "vec" + std::to_string(n) + " needs a component type here");
                out = vec_t(n, T(Type::F32)); if (inferred) *inferred = true; return true;
            }
            if (s.size() == 5) {
                Type e;
                switch (s[4]) { case 'f': e = T(Type::F32); break; case 'i': e = T(Type::I32); break; case 'u': e = T(Type::U32); break;
                                case 'h': e = T(Type::F16); break; default: return false; }
                p++; out = vec_t(n, e); return true;
            }
            return false;
        }
        if (s.size() >= 6 && s.compare(0, 3, "mat") == 0 && isdigit((unsigned char)s[3]) && s[4] == 'x' && isdigit((unsigned char)s[5])) {
            const int c = s[3] - '0', r = s[5] - '0';
            if (s.size() == 6) {
                p++;
                Type e;
                if (templ_scalar(e)) { if (e.k == Type::F16) err(line, "f16 matrices are not supported"); out = mat_t(c, r, e); return true; }
                if (!allow_infer) err(line, "matrix type needs a component type here");
                out = mat_t(c, r, T(Type::F32)); if (inferred) *inferred = true; return true;
            }
            if (s.size() == 7 && s[6] == 'h') err(line, "f16 matrices are not supported");
            if (s.size() == 7 && s[6] == 'f') { p++; out = mat_t(c, r, T(Type::F32)); return true; }
            return false;
        }
        if (s == "array") {
            p++;
            if (!accept("<")) { if (!allow_infer) err(line, "array needs template arguments here"); out = array_t(T(Type::F32), 0); if (inferred) *inferred = true; return true; }
            Type e; if (!try_type(e)) err(line, "expected array element type");
            int count = 0;
            if (accept(",")) {
                if (cur().k == Tok::Int) { count = (int)cur().f; p++; }
                else {
                    const std::string cn = ident();
                    bool found = false;
                    for (auto& g : m.globals) if (g.name == cn && g.space == "const" && g.init && g.init->k == Expr::Lit) { count = (int)g.init->f; found = true; }
                    if (!found) err(line, "array length must be an integer literal or a literal const");
                }
            }
            close_angle();
            out = array_t(e, count);
            return true;
        }
        if (s == "texture_2d") { p++; expect("<"); Type e; try_type(e); close_angle(); if (e.k != Type::F32) err(line, "only texture_2d<f32> is supported"); out = T(Type::Texture2D); return true; }
        if (s == "sampler") { p++; out = T(Type::Sampler); return true; }
        static const std::set<std::string> unsupported = {"ptr", "atomic", "sampler_comparison", "texture_1d", "texture_3d", "texture_cube",
            "texture_cube_array", "texture_2d_array", "texture_multisampled_2d", "texture_depth_2d", "texture_depth_2d_array",
            "texture_depth_cube", "texture_depth_cube_array", "texture_depth_multisampled_2d", "texture_storage_1d",
            "texture_storage_2d", "texture_storage_2d_array", "texture_storage_3d", "texture_external"};
        if (unsupported.count(s)) err(line, "type '" + s + "' is not supported");
        auto al = m.aliases.find(s);
        if (al != m.aliases.end()) { p++; out = al->second; return true; }
        for (auto& sd : m.structs) if (sd->name == s) { p++; out = T(Type::Struct); out.st = sd.get(); return true; }
        return false;
    }
    void close_angle() {
        if (is(">")) { p++; return; }
        if (is(">>")) { t[p].s = ">"; return; }        // split '>>' closing two templates
        if (is(">=")) { t[p].s = "="; return; }
        err(cur().line, "expected '>'");
    }
    Type type() { Type x; if (!try_type(x)) err(cur().line, "expected a type, found '" + cur().s + "'"); return x; }

    // ---- expressions ----
    ExprP mk(Expr::K k, int line) { auto e = std::make_shared<Expr>(); e->k = k; e->line = line; return e; }
    ExprP primary() {
        const Tok& c = cur();
        const int line = c.line;
        if (c.k == Tok::Int || c.k == Tok::Float) {
            auto e = mk(Expr::Lit, line);
            e->f = c.f; e->is_float = c.k == Tok::Float; e->suffix = c.suffix;
            p++; return e;
        }
        if (is("true") || is("false")) { auto e = mk(Expr::Lit, line); e->is_bool = true; e->f = is("true") ? 1 : 0; p++; return e; }
        if (accept("(")) { auto e = expr(); expect(")"); return e; }
        if (c.k == Tok::Ident) {
            if (c.s == "bitcast") {
                p++; expect("<"); Type to = type(); close_angle(); expect("(");
                auto e = mk(Expr::Call, line); e->name = "bitcast"; e->ctor = to; e->args.push_back(expr()); expect(")"); return e;
            }
            // type constructor / conversion?
            const size_t save = p;
            Type ty; bool inferred = false;
            if (try_type(ty, true, &inferred)) {
                if (is("(")) {
                    auto e = mk(Expr::Construct, line);
                    e->ctor = ty; e->ctor_infer = inferred;
                    p++;
                    while (!is(")")) { e->args.push_back(expr()); if (!accept(",")) break; }
                    expect(")");
                    return e;
                }
                p = save;
                err(line, "type '" + c.s + "' used as a value");
            }
            const std::string name = ident();
            if (is("(")) {
                auto e = mk(Expr::Call, line); e->name = name;
                p++;
                while (!is(")")) { e->args.push_back(expr()); if (!accept(",")) break; }
                expect(")");
                return e;
            }
            auto e = mk(Expr::Ident, line); e->name = name; return e;
        }
        err(line, "unexpected token '" + c.s + "' in expression");
    }
    ExprP postfix() {
        auto e = primary();
        for (;;) {
            const int line = cur().line;
            if (accept(".")) { auto m2 = mk(Expr::Member, line); m2->name = ident(); m2->args.push_back(e); e = m2; }
            else if (accept("[")) { auto x = mk(Expr::Index, line); x->args.push_back(e); x->args.push_back(expr()); expect("]"); e = x; }
            else break;
        }
        return e;
    }
    ExprP unary() {
        const int line = cur().line;
        for (const char* op : {"-", "!", "~"})
            if (is(op)) { p++; auto e = mk(Expr::Unary, line); e->name = op; e->args.push_back(unary()); return e; }
        if (is("&")) { p++; auto e = mk(Expr::Unary, line); e->name = "&"; e->args.push_back(unary()); return e; }   // arrayLength(&buffer.array) only
        if (is("*")) err(line, "pointers are not supported");
        return postfix();
    }
    ExprP binary_level(int level) {
        static const std::vector<std::vector<std::string>> levels = {
            {"||"}, {"&&"}, {"|"}, {"^"}, {"&"}, {"==", "!="}, {"<", ">", "<=", ">="}, {"<<", ">>"}, {"+", "-"}, {"*", "/", "%"}};
        if (level == (int)levels.size()) return unary();
        auto lhs = binary_level(level + 1);
        for (;;) {
            bool found = false;
            for (auto& op : levels[level]) {
                if (cur().k == Tok::Punct && cur().s == op) {
                    const int line = cur().line;
                    p++;
                    auto rhs = binary_level(level + 1);
                    auto e = mk(Expr::Binary, line); e->name = op; e->args = {lhs, rhs};
                    lhs = e; found = true; break;
                }
            }
            if (!found) return lhs;
        }
    }
    ExprP expr() { return binary_level(0); }

    // ---- statements ----
    std::vector<StmtP> block() {
        expect("{");
        std::vector<StmtP> out;
        while (!is("}")) { if (cur().k == Tok::End) err(cur().line, "unterminated block"); auto s = statement(); if (s) out.push_back(s); }
        expect("}");
        return out;
    }
    StmtP mks(Stmt::K k, int line) { auto s = std::make_shared<Stmt>(); s->k = k; s->line = line; return s; }
    StmtP simple_statement() {   // without the trailing ';' (used by for-headers)
        const int line = cur().line;
        if (is("let") || is("var") || is("const")) {
            auto s = mks(is("let") ? Stmt::Let : is("var") ? Stmt::Var : Stmt::Const, line);
            p++;
            if (s->k == Stmt::Var && accept("<")) { ident(); close_angle(); }   // var<function>
            s->name = ident();
            if (accept(":")) { s->has_type = true; s->decl_type = type(); }
            if (accept("=")) s->a = expr();
            if (s->k != Stmt::Var && !s->a) err(line, "declaration needs an initialiser");
            return s;
        }
        if (is("_")) { p++; expect("="); auto s = mks(Stmt::CallS, line); s->a = expr(); return s; }
        auto lhs = expr();
        if (cur().k == Tok::Punct) {
            static const char* ops[] = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", nullptr};
            for (int i = 0; ops[i]; i++)
                if (cur().s == ops[i]) { p++; auto s = mks(Stmt::Assign, line); s->op = ops[i]; s->a = lhs; s->b = expr(); return s; }
            if (cur().s == "++" || cur().s == "--") { auto s = mks(Stmt::Incr, line); s->op = cur().s; s->a = lhs; p++; return s; }
        }
        if (lhs->k != Expr::Call) err(line, "expression statement must be a function call");
        auto s = mks(Stmt::CallS, line); s->a = lhs; return s;
    }
    StmtP statement() {
        const int line = cur().line;
        if (accept(";")) return nullptr;
        if (is("{")) { auto s = mks(Stmt::Block, line); s->body = block(); return s; }
        if (accept("if")) {
            auto s = mks(Stmt::If, line);
            s->a = expr();
            s->body = block();
            if (accept("else")) {
                if (is("if")) { s->else_body.push_back(statement()); }
                else s->else_body = block();
            }
            return s;
        }
        if (accept("for")) {
            auto s = mks(Stmt::For, line);
            expect("(");
            if (!is(";")) s->init = simple_statement();
            expect(";");
            if (!is(";")) s->a = expr();
            expect(";");
            if (!is(")")) s->update = simple_statement();
            expect(")");
            s->body = block();
            return s;
        }
        if (accept("while")) { auto s = mks(Stmt::While, line); s->a = expr(); s->body = block(); return s; }
        if (accept("loop")) {
            auto s = mks(Stmt::Loop, line);
            expect("{");
            while (!is("}")) {
                if (accept("continuing")) {
                    expect("{");
                    while (!is("}")) {
                        if (is("break") && peek().k == Tok::Ident && peek().s == "if") {
                            const int l2 = cur().line; p += 2;
                            auto b = mks(Stmt::BreakIf, l2); b->a = expr(); expect(";");
                            s->continuing.push_back(b);
                        } else { auto c = statement(); if (c) s->continuing.push_back(c); }
                    }
                    expect("}");
                } else { auto c = statement(); if (c) s->body.push_back(c); }
            }
            expect("}");
            return s;
        }
        if (accept("switch")) {
            auto s = mks(Stmt::Switch, line);
            s->a = expr();
            expect("{");
            while (!is("}")) {
                Stmt::Case c;
                if (accept("default")) { c.is_default = true; }
                else {
                    expect("case");
                    for (;;) {
                        if (accept("default")) c.is_default = true; else c.sel.push_back(expr());
                        if (!accept(",")) break;
                        if (is(":") || is("{")) break;
                    }
                }
                accept(":");
                c.body = block();
                s->cases.push_back(c);
            }
            expect("}");
            return s;
        }
        if (accept("break")) { expect(";"); return mks(Stmt::Break, line); }
        if (accept("continue")) { expect(";"); return mks(Stmt::Continue, line); }
        if (accept("discard")) { expect(";"); return mks(Stmt::Discard, line); }
        if (accept("return")) { auto s = mks(Stmt::Return, line); if (!is(";")) s->a = expr(); expect(";"); return s; }
        auto s = simple_statement();
        expect(";");
        return s;
    }

    // ---- module scope ----
    static int attr_int(const std::vector<Attr>& as, const char* name, int dflt) {
        for (auto& a : as) if (a.name == name && !a.args.empty()) return atoi(a.args[0].c_str());
        return dflt;
    }
    void parse_module() {
        while (cur().k != Tok::End) {
            const int line = cur().line;
            if (accept(";")) continue;
            if (is("enable") || is("requires") || is("diagnostic")) { while (!is(";")) p++; p++; continue; }
            std::vector<Attr> as = attrs();
            if (accept("struct")) {
                auto sd = std::make_shared<StructDecl>();
                sd->name = ident();
                expect("{");
                while (!is("}")) {
                    Member mb; mb.line = cur().line;
                    mb.attrs = attrs();
                    mb.name = ident(); expect(":"); mb.type = type();
                    sd->members.push_back(mb);
                    if (!accept(",")) break;
                }
                expect("}");
                accept(";");
                m.structs.push_back(sd);
                m.decl_order.push_back("s:" + sd->name);
                continue;
            }
            if (accept("alias")) { const std::string n = ident(); expect("="); m.aliases[n] = type(); expect(";"); continue; }
            if (is("var") || is("const") || is("override")) {
                Global g; g.line = line;
                const std::string kw = cur().s; p++;
                g.space = kw == "const" ? "const" : "handle";
                if (kw == "override") err(line, "override declarations are not supported (expression/override.rs is a todo!() in the reference)");
                if (kw == "var" && accept("<")) {
                    g.space = ident();
                    if (accept(",")) ident();
                    close_angle();
                    if (g.space == "workgroup") err(line, "workgroup memory is not supported");
                }
                g.name = ident();
                bool has_type = false;
                if (accept(":")) { g.type = type(); has_type = true; }
                if (accept("=")) g.init = expr();
                expect(";");
                if (!has_type && !g.init) err(line, "global needs a type or an initialiser");
                if (!has_type) g.type = T(Type::Void);   // inferred by the checker
                g.group = attr_int(as, "group", -1);
                g.binding = attr_int(as, "binding", -1);
                m.globals.push_back(g);
                m.decl_order.push_back("g:" + g.name);
                continue;
            }
            if (accept("fn")) {
                auto f = std::make_shared<Function>();
                f->line = line; f->attrs = as;
                f->name = ident();
                expect("(");
                while (!is(")")) {
                    Param pr; pr.line = cur().line;
                    pr.attrs = attrs();
                    pr.name = ident(); expect(":"); pr.type = type();
                    f->params.push_back(pr);
                    if (!accept(",")) break;
                }
                expect(")");
                f->ret = T(Type::Void);
                if (accept("->")) { f->ret_attrs = attrs(); f->ret = type(); }
                f->body = block();
                m.functions.push_back(f);
                m.decl_order.push_back("f:" + f->name);
                continue;
            }
            err(line, "unexpected token '" + cur().s + "' at module scope");
        }
    }
};

// ------------------------------------------------------------------------------------------
// checker + emitter
// ------------------------------------------------------------------------------------------
std::string fmt_float(double v) {
    if (std::isinf(v) || std::isnan(v)) throw std::runtime_error("WGSL: non-finite float literal");
    char buf[64];
    snprintf(buf, sizeof(buf), "%.9g", (double)(float)v);
    std::string s = buf;
    if (s.find_first_of(".en") == std::string::npos) s += ".0";
    return s + "f";
}

struct Emitter {
    Module& m;
    int stage;
    std::string entry;
    std::string ns;
    std::string out;
    // scopes
    struct Var { Type type; bool is_const; std::string cname; };
    std::vector<std::map<std::string, Var>> scopes;
    const Function* cur_fn = nullptr;
    int label_counter = 0;
    std::vector<std::string> continue_labels;   // innermost `loop` continuing target ("" for for/while)
    std::map<std::string, const Function*> fn_by_name;
    std::map<std::string, const Global*> global_by_name;
    bool has_private = false;

    Emitter(Module& mod, int st, const std::string& e) : m(mod), stage(st), entry(e) { ns = st == 1 ? "wgb_vertex" : "wgb_fragment"; }

    // ---- struct layout ----
    void layout_struct(StructDecl& sd, int line) {
        uint32_t off = 0, maxa = 4;
        bool shareable = true;
        for (auto& mb : sd.members) {
            if (mb.type.k == Type::Bool || (mb.type.k == Type::Vec && mb.type.elem->k == Type::Bool)) { shareable = false; break; }
        }
        if (!shareable) { sd.size = 0; sd.align = 0; return; }
        for (auto& mb : sd.members) {
            uint32_t a, s;
            layout_of(mb.type, a, s, mb.line);
            for (auto& at : mb.attrs) { if (at.name == "align" && !at.args.empty()) a = (uint32_t)atoi(at.args[0].c_str()); }
            off = (off + a - 1) / a * a;
            mb.offset = off;
            for (auto& at : mb.attrs) { if (at.name == "size" && !at.args.empty()) s = (uint32_t)atoi(at.args[0].c_str()); }
            off += s;
            maxa = std::max(maxa, a);
        }
        sd.align = maxa;
        sd.size = (off + maxa - 1) / maxa * maxa;
        (void)line;
    }

    // ---- scopes ----
    void push() { scopes.emplace_back(); }
    void pop() { scopes.pop_back(); }
    void declare(const std::string& n, const Type& t, bool is_const, int line) {
        if (scopes.back().count(n)) err(line, "redeclaration of '" + n + "'");
        scopes.back()[n] = Var{t, is_const, n};
    }
    const Var* lookup(const std::string& n) const {
        for (auto it = scopes.rbegin(); it != scopes.rend(); ++it) { auto f = it->find(n); if (f != it->end()) return &f->second; }
        return nullptr;
    }

    // ---- abstract numerics ----
    static Type concretize(const Type& t) {
        if (t.k == Type::AInt) return T(Type::I32);
        if (t.k == Type::AFloat) return T(Type::F32);
        if ((t.k == Type::Vec || t.k == Type::Mat || t.k == Type::Array) && t.elem && t.elem->is_abstract()) {
            Type r = t; r.elem = std::make_shared<Type>(concretize(*t.elem)); return r;
        }
        return t;
    }
    // can a value of type `from` be used where `to` is expected (abstract conversions only)?
    static bool converts(const Type& from, const Type& to) {
        if (same(from, to)) return true;
        if (from.k == Type::AInt) return to.k == Type::I32 || to.k == Type::U32 || to.k == Type::F32 || to.k == Type::F16 || to.k == Type::AFloat;
        if (from.k == Type::AFloat) return to.k == Type::F32 || to.k == Type::F16;
        if (from.k == Type::Vec && to.k == Type::Vec && from.n == to.n) return converts(*from.elem, *to.elem);
        if (from.k == Type::Mat && to.k == Type::Mat && from.n == to.n && from.rows == to.rows) return converts(*from.elem, *to.elem);
        if (from.k == Type::Array && to.k == Type::Array && from.count == to.count) return converts(*from.elem, *to.elem);
        return false;
    }

    // ---- expression checking + emission: returns CUDA text, sets e.type ----
    struct Val { std::string s; Type t; bool is_const_num = false; double num = 0; };

    std::string lit_text(const Val& v, const Type& want, int line) {
        // a folded abstract constant rendered as `want`
        if (want.k == Type::F32 || want.k == Type::AFloat) return fmt_float(v.num);
        if (want.k == Type::F16) {          // the literal rounded to binary16 (65504 is the largest finite value; halfway cases round to even)
            if (std::fabs(v.num) >= 65520.0) err(line, "literal out of range for f16");
            return "wgb_f16(" + fmt_float((double)(float)v.num) + ")";
        }
        if (want.k == Type::U32) { if (v.num < 0 || v.num > 4294967295.0) err(line, "literal out of range for u32"); return std::to_string((unsigned long long)v.num) + "u"; }
        if (want.k == Type::I32 || want.k == Type::AInt) {
            if (v.num < -2147483648.0 || v.num > 2147483647.0) err(line, "literal out of range for i32");
            if (v.num == -2147483648.0) return "(-2147483647 - 1)";
            return std::to_string((long long)v.num);
        }
        err(line, "cannot convert a numeric literal to " + wgsl_type_name(want));
    }
    // coerce a value to a concrete type (abstract -> concrete); error if impossible
    Val coerce(Val v, const Type& to, int line) {
        if (same(v.t, to)) return v;
        if (!converts(v.t, to)) err(line, "cannot convert " + wgsl_type_name(v.t) + " to " + wgsl_type_name(to));
        if (v.t.is_scalar()) {
            if (v.is_const_num) { v.s = lit_text(v, to, line); v.t = to; return v; }
            // non-constant abstract cannot exist
            v.t = to; return v;
        }
        // abstract vector/matrix/array: built from wgb constructors with int/float literals -> re-tag
        if (v.t.k == Type::Vec) { v.s = cuda_type(to, line) + "(" + v.s + ")"; v.t = to; return v; }
        v.t = to;
        return v;
    }
    Val concrete(Val v, int line) { return coerce(v, concretize(v.t), line); }

    Val expr(const ExprP& e) {
        Val v = expr_inner(e);
        e->type = v.t;
        return v;
    }

    static bool is_swizzle(const std::string& s, int n, std::vector<int>& idx) {
        if (s.empty() || s.size() > 4) return false;
        const char* sets[2] = {"xyzw", "rgba"};
        for (int k = 0; k < 2; k++) {
            idx.clear();
            bool ok = true;
            for (char c : s) { const char* q = strchr(sets[k], c); if (!q || (q - sets[k]) >= n) { ok = false; break; } idx.push_back((int)(q - sets[k])); }
            if (ok) return true;
        }
        return false;
    }

    // resource access chains rooted at a uniform/storage global: returns true and fills the
    // byte-offset expression (constant part + dynamic part) and the accessed type
    bool resource_chain(const ExprP& e, const Global*& g, uint32_t& const_off, std::string& dyn_off, Type& ty) {
        if (e->k == Expr::Ident) {
            if (lookup(e->name)) return false;
            auto it = global_by_name.find(e->name);
            if (it == global_by_name.end()) return false;
            if (it->second->space != "uniform" && it->second->space != "storage") return false;
            g = it->second; const_off = 0; dyn_off.clear(); ty = g->type;
            return true;
        }
        if (e->k == Expr::Member) {
            Type base;
            if (!resource_chain(e->args[0], g, const_off, dyn_off, base)) return false;
            if (base.k == Type::Struct) {
                for (auto& mb : base.st->members) if (mb.name == e->name) { const_off += mb.offset; ty = mb.type; return true; }
                err(e->line, "struct " + base.st->name + " has no member '" + e->name + "'");
            }
            if (base.k == Type::Vec) {
                std::vector<int> idx;
                if (is_swizzle(e->name, base.n, idx) && idx.size() == 1) { const_off += 4 * idx[0]; ty = *base.elem; return true; }
                return false;   // multi-component swizzle: load the vector, swizzle by value
            }
            return false;
        }
        if (e->k == Expr::Index) {
            Type base;
            if (!resource_chain(e->args[0], g, const_off, dyn_off, base)) return false;
            uint32_t stride; Type elem;
            if (base.k == Type::Array) { uint32_t a, s; layout_of(*base.elem, a, s, e->line); stride = (s + a - 1) / a * a; elem = *base.elem; }
            else if (base.k == Type::Mat) { stride = base.rows == 2 ? 8 : 16; elem = vec_t(base.rows, *base.elem); }
            else if (base.k == Type::Vec) { stride = 4; elem = *base.elem; }
            else err(e->line, "cannot index " + wgsl_type_name(base));
            Val iv = expr(e->args[1]);
            if (iv.is_const_num) { const_off += (uint32_t)iv.num * stride; }
            else {
                iv = concrete(iv, e->line);
                if (iv.t.k != Type::I32 && iv.t.k != Type::U32) err(e->line, "index must be an integer");
                // out-of-range dynamic indices are clamped to the last element (the reference aborts with PointerOutOfBounds)
                std::string idx = "(u32)(" + iv.s + ")";
                if (base.k != Type::Array || base.count) {
                    const int cnt = base.k == Type::Array ? base.count : base.n;
                    idx = "min(" + idx + ", " + std::to_string(cnt - 1) + "u)";
                }
                dyn_off += (dyn_off.empty() ? "" : " + ") + idx + " * " + std::to_string(stride) + "u";
            }
            ty = elem;
            return true;
        }
        return false;
    }
    std::string emit_load(const Global* g, const Type& ty, uint32_t const_off, const std::string& dyn, int line) {
        const std::string off = std::to_string(const_off) + "u" + (dyn.empty() ? "" : " + " + dyn);
        const std::string gb = "wgb, " + std::to_string(g->group) + ", " + std::to_string(g->binding) + ", ";
        switch (ty.k) {
            case Type::F32: case Type::I32: case Type::U32: case Type::Vec: case Type::Mat:
                return "wgb_load<" + cuda_type(ty, line) + ">(" + gb + off + ")";
            case Type::Struct: {
                std::string s = ty.st->name + "{";
                bool first = true;
                for (auto& mb : ty.st->members) {
                    if (!first) s += ", ";
                    first = false;
                    s += emit_load(g, mb.type, const_off + mb.offset, dyn, line);
                }
                return s + "}";
            }
            case Type::Array: {
                if (ty.count == 0) err(line, "a runtime-sized array cannot be loaded by value");
                uint32_t a, sz; layout_of(*ty.elem, a, sz, line);
                const uint32_t stride = (sz + a - 1) / a * a;
                std::string s = cuda_type(ty, line) + "{{";
                for (int i = 0; i < ty.count; i++) { if (i) s += ", "; s += emit_load(g, *ty.elem, const_off + i * stride, dyn, line); }
                return s + "}}";
            }
            default: err(line, "cannot load " + wgsl_type_name(ty) + " from a buffer");
        }
    }

    Val expr_inner(const ExprP& e) {
        const int line = e->line;
        switch (e->k) {
            case Expr::Lit: {
                Val v;
                if (e->is_bool) { v.s = e->f != 0 ? "true" : "false"; v.t = T(Type::Bool); return v; }
                v.is_const_num = true; v.num = e->f;
                if (e->is_float || e->suffix == 'h') v.t = e->suffix == 'f' ? T(Type::F32) : e->suffix == 'h' ? T(Type::F16) : T(Type::AFloat);
                else v.t = e->suffix == 'u' ? T(Type::U32) : e->suffix == 'i' ? T(Type::I32) : T(Type::AInt);
                v.s = lit_text(v, v.t, line);
                return v;
            }
            case Expr::Ident: {
                if (const Var* var = lookup(e->name)) { Val v; v.s = var->cname; v.t = var->type; return v; }
                auto gi = global_by_name.find(e->name);
                if (gi != global_by_name.end()) {
                    const Global* g = gi->second;
                    Val v; v.t = g->type;
                    if (g->space == "uniform" || g->space == "storage") { v.s = emit_load(g, g->type, 0, "", line); return v; }
                    if (g->space == "private") { v.s = "wgb_inv." + g->name; return v; }
                    if (g->space == "const") { v.s = g->type.is_scalar() ? g->name : g->name + "()"; return v; }
                    err(line, "'" + e->name + "' is a resource handle and can only be passed to texture builtins");
                }
                err(line, "unknown identifier '" + e->name + "'");
            }
            case Expr::Unary: {
                if (e->name == "&") err(line, "pointers are not supported (except arrayLength(&runtime_sized_array))");
                Val a = expr(e->args[0]);
                Val v; v.t = a.t;
                if (a.is_const_num && e->name == "-") { v.is_const_num = true; v.num = -a.num; v.s = lit_text(v, v.t, line); return v; }
                a = concrete(a, line); v.t = a.t;
                if (e->name == "-") {
                    const Type& sc = a.t.scalar();
                    if (!(sc.k == Type::F32 || sc.k == Type::F16 || sc.k == Type::I32)) err(line, "unary '-' needs a signed numeric operand");
                    v.s = "(-" + a.s + ")";
                } else if (e->name == "!") {
                    if (a.t.scalar().k != Type::Bool) err(line, "'!' needs a bool operand");
                    v.s = a.t.k == Type::Vec ? "wgb_not(" + a.s + ")" : "(!" + a.s + ")";
                } else {
                    if (!(a.t.scalar().k == Type::I32 || a.t.scalar().k == Type::U32)) err(line, "'~' needs an integer operand");
                    v.s = a.t.k == Type::Vec ? "wgb_bitnot(" + a.s + ")" : "(~" + a.s + ")";
                }
                return v;
            }
            case Expr::Binary: return binary(e);
            case Expr::Member: {
                const Global* g; uint32_t co; std::string dyn; Type ty;
                if (resource_chain(e, g, co, dyn, ty)) { Val v; v.s = emit_load(g, ty, co, dyn, line); v.t = ty; return v; }
                Val a = expr(e->args[0]);
                if (a.t.k == Type::Struct) {
                    for (auto& mb : a.t.st->members) if (mb.name == e->name) { Val v; v.s = a.s + "." + mb.name; v.t = mb.type; return v; }
                    err(line, "struct " + a.t.st->name + " has no member '" + e->name + "'");
                }
                if (a.t.k == Type::Vec) {
                    std::vector<int> idx;
                    if (!is_swizzle(e->name, a.t.n, idx)) err(line, "invalid swizzle '." + e->name + "' on " + wgsl_type_name(a.t));
                    static const char* comp = "xyzw";
                    a = concrete(a, line);
                    Val v;
                    if (idx.size() == 1) { v.s = a.s + "." + comp[idx[0]]; v.t = *a.t.elem; return v; }
                    v.t = vec_t((int)idx.size(), *a.t.elem);
                    std::string args;
                    for (size_t i = 0; i < idx.size(); i++) { if (i) args += ", "; args += std::string("wgb_t.") + comp[idx[i]]; }
                    v.s = "([&] { const auto wgb_t = " + a.s + "; return " + cuda_type(v.t, line) + "(" + args + "); }())";
                    return v;
                }
                err(line, "cannot access member '" + e->name + "' of " + wgsl_type_name(a.t));
            }
            case Expr::Index: {
                const Global* g; uint32_t co; std::string dyn; Type ty;
                if (resource_chain(e, g, co, dyn, ty)) { Val v; v.s = emit_load(g, ty, co, dyn, line); v.t = ty; return v; }
                Val a = expr(e->args[0]);
                Val i = expr(e->args[1]);
                if (!i.is_const_num) i = concrete(i, line);
                if (!(i.t.k == Type::I32 || i.t.k == Type::U32 || i.t.k == Type::AInt)) err(line, "index must be an integer");
                a = concrete(a, line);
                Val v;
                int cnt;
                if (a.t.k == Type::Vec) { v.t = *a.t.elem; cnt = a.t.n; }
                else if (a.t.k == Type::Mat) { v.t = vec_t(a.t.rows, *a.t.elem); cnt = a.t.n; }
                else if (a.t.k == Type::Array) { v.t = *a.t.elem; cnt = a.t.count; }
                else err(line, "cannot index " + wgsl_type_name(a.t));
                if (i.is_const_num) {
                    if (i.num < 0 || i.num >= cnt) err(line, "index out of bounds");
                    v.s = a.s + "[" + std::to_string((int)i.num) + "]";
                } else v.s = a.s + "[min((u32)(" + i.s + "), " + std::to_string(cnt - 1) + "u)]";
                return v;
            }
            case Expr::Construct: return construct(e);
            case Expr::Call: return call(e);
        }
        err(line, "internal: unknown expression kind");
    }

    Val binary(const ExprP& e) {
        const int line = e->line;
        const std::string& op = e->name;
        Val a = expr(e->args[0]), b = expr(e->args[1]);
        const bool arith = op == "+" || op == "-" || op == "*" || op == "/" || op == "%";
        const bool cmp = op == "==" || op == "!=" || op == "<" || op == ">" || op == "<=" || op == ">=";
        // constant folding of abstract (x) abstract, in f64 / i64 like naga's constant evaluator
        if (a.is_const_num && b.is_const_num && a.t.is_abstract() && b.t.is_abstract() && (arith || cmp)) {
            const bool fl = a.t.k == Type::AFloat || b.t.k == Type::AFloat;
            Val v;
            if (cmp) {
                bool r = op == "==" ? a.num == b.num : op == "!=" ? a.num != b.num : op == "<" ? a.num < b.num : op == ">" ? a.num > b.num : op == "<=" ? a.num <= b.num : a.num >= b.num;
                v.s = r ? "true" : "false"; v.t = T(Type::Bool); return v;
            }
            v.is_const_num = true; v.t = fl ? T(Type::AFloat) : T(Type::AInt);
            if (op == "+") v.num = a.num + b.num; else if (op == "-") v.num = a.num - b.num; else if (op == "*") v.num = a.num * b.num;
            else if (op == "/") { if (b.num == 0) err(line, "division by zero in a constant expression"); v.num = fl ? a.num / b.num : (double)((long long)a.num / (long long)b.num); }
            else { if (b.num == 0) err(line, "remainder by zero in a constant expression"); v.num = fl ? std::fmod(a.num, b.num) : (double)((long long)a.num % (long long)b.num); }
            v.s = lit_text(v, v.t, line);
            return v;
        }
        // unify abstract operands with the concrete side
        auto unify_scalar = [&](Val& x, const Type& other) {
            if (!x.t.is_abstract()) return;
            Type target = other.scalar();
            if (target.is_abstract()) target = concretize((x.t.scalar().k == Type::AFloat || target.k == Type::AFloat) ? T(Type::AFloat) : T(Type::AInt));
            if (x.t.scalar().k == Type::AFloat && !(target.k == Type::F32)) target = T(Type::F32);
            if (x.t.is_scalar()) x = coerce(x, target, line);
            else if (x.t.k == Type::Vec) x = coerce(x, vec_t(x.t.n, target), line);
            else if (x.t.k == Type::Mat) x = coerce(x, mat_t(x.t.n, x.t.rows, target), line);
        };
        unify_scalar(a, b.t);
        unify_scalar(b, a.t);
        Val v;
        if (op == "&&" || op == "||") {
            if (a.t.k != Type::Bool || b.t.k != Type::Bool) err(line, "'" + op + "' needs bool operands");
            v.s = "(" + a.s + " " + op + " " + b.s + ")"; v.t = T(Type::Bool); return v;
        }
        if (cmp) {
            if (!same(a.t, b.t)) err(line, "comparison of " + wgsl_type_name(a.t) + " with " + wgsl_type_name(b.t));
            if (a.t.is_scalar()) { v.s = "(" + a.s + " " + op + " " + b.s + ")"; v.t = T(Type::Bool); return v; }
            if (a.t.k == Type::Vec) {
                static const std::map<std::string, std::string> names = {{"==", "wgb_eq"}, {"!=", "wgb_ne"}, {"<", "wgb_lt"}, {">", "wgb_gt"}, {"<=", "wgb_le"}, {">=", "wgb_ge"}};
                v.s = names.at(op) + "(" + a.s + ", " + b.s + ")"; v.t = vec_t(a.t.n, T(Type::Bool)); return v;
            }
            err(line, "cannot compare " + wgsl_type_name(a.t));
        }
        if (op == "<<" || op == ">>") {
            const Type& sa = a.t.scalar();
            if (!(sa.k == Type::I32 || sa.k == Type::U32)) err(line, "shift needs an integer left operand");
            if (b.t.scalar().k == Type::I32 && b.is_const_num) b = coerce(b, T(Type::U32), line);
            if (a.t.k == Type::Vec) { v.s = (op == "<<" ? "wgb_shl(" : "wgb_shr(") + a.s + ", " + b.s + ")"; }
            else v.s = "(" + a.s + " " + op + " ((u32)(" + b.s + ") & 31u))";
            v.t = a.t; return v;
        }
        if (op == "&" || op == "|" || op == "^") {
            if (!same(a.t, b.t)) err(line, "'" + op + "' needs operands of the same type");
            const Type& sa = a.t.scalar();
            if (!(sa.k == Type::I32 || sa.k == Type::U32 || sa.k == Type::Bool)) err(line, "'" + op + "' needs integer or bool operands");
            if (a.t.k == Type::Vec) { v.s = std::string(op == "&" ? "wgb_and(" : op == "|" ? "wgb_or(" : "wgb_xor(") + a.s + ", " + b.s + ")"; }
            else v.s = "(" + a.s + " " + op + " " + b.s + ")";
            v.t = a.t; return v;
        }
        if (!arith) err(line, "unknown operator '" + op + "'");
        // result type
        const Type &ta = a.t, &tb = b.t;
        if (ta.scalar().k == Type::Bool || tb.scalar().k == Type::Bool) err(line, "arithmetic on bool");
        if (!same(ta.scalar(), tb.scalar())) err(line, "operands of '" + op + "' have different component types: " + wgsl_type_name(ta) + " and " + wgsl_type_name(tb));
        const bool fl = ta.scalar().k == Type::F32 || ta.scalar().k == Type::F16;
        if (ta.is_scalar() && tb.is_scalar()) {
            v.t = ta;
            if (fl) {
                static const std::map<std::string, std::string> fn = {{"+", "wgb_add"}, {"-", "wgb_sub"}, {"*", "wgb_mul"}, {"/", "wgb_div"}, {"%", "wgb_rem"}};
                v.s = fn.at(op) + "(" + a.s + ", " + b.s + ")";
            } else if (op == "/") v.s = "wgb_idiv(" + a.s + ", " + b.s + ")";
            else if (op == "%") v.s = "wgb_irem(" + a.s + ", " + b.s + ")";
            else v.s = "(" + a.s + " " + op + " " + b.s + ")";
            return v;
        }
        if (ta.k == Type::Mat || tb.k == Type::Mat) {
            if (op == "*") {
                if (ta.k == Type::Mat && tb.k == Type::Vec) { if (ta.n != tb.n) err(line, "matrix * vector dimension mismatch"); v.t = vec_t(ta.rows, *ta.elem); }
                else if (ta.k == Type::Mat && tb.k == Type::Mat) { if (ta.n != tb.rows) err(line, "matrix * matrix dimension mismatch"); v.t = mat_t(tb.n, ta.rows, *ta.elem); }
                else if (ta.k == Type::Mat && tb.is_scalar()) v.t = ta;
                else if (ta.is_scalar() && tb.k == Type::Mat) v.t = tb;
                else err(line, "vector * matrix is not supported (binary.rs:266 is a todo!() in the reference)");
                v.s = "(" + a.s + " * " + b.s + ")";
                return v;
            }
            if ((op == "+" || op == "-") && same(ta, tb)) { v.t = ta; v.s = "(" + a.s + " " + op + " " + b.s + ")"; return v; }
            err(line, "unsupported matrix operation '" + op + "'");
        }
        // vector (x) vector, vector (x) scalar, scalar (x) vector
        if (ta.k == Type::Vec && tb.k == Type::Vec && ta.n != tb.n) err(line, "vector size mismatch");
        v.t = ta.k == Type::Vec ? ta : tb;
        if (fl && op == "%") v.s = "wgb_rem(" + a.s + ", " + b.s + ")";
        else if (!fl && op == "/") v.s = "wgb_idiv(" + a.s + ", " + b.s + ")";
        else if (!fl && op == "%") v.s = "wgb_irem(" + a.s + ", " + b.s + ")";
        else v.s = "(" + a.s + " " + op + " " + b.s + ")";
        return v;
    }

    Val construct(const ExprP& e) {
        const int line = e->line;
        Type to = e->ctor;
        std::vector<Val> args;
        for (auto& a : e->args) args.push_back(expr(a));
        Val v;
        if (to.is_scalar()) {   // conversion (expression/as.rs:60-184)
            if (args.size() == 0) { v.t = to; v.s = cuda_type(to, line) + "(0)"; return v; }
            if (args.size() != 1) err(line, "scalar conversion takes one argument");
            Val a = args[0];
            if (a.is_const_num && a.t.is_abstract()) {
                if (to.k == Type::Bool) { v.s = a.num != 0 ? "true" : "false"; v.t = to; return v; }
                Val c = a; if (to.k != Type::F32 && to.k != Type::F16 && a.t.k == Type::AFloat) c.num = std::trunc(a.num);
                c.t = to; c.s = lit_text(c, to, line); c.is_const_num = true; return c;
            }
            a = concrete(a, line);
            if (!a.t.is_scalar()) err(line, "cannot convert " + wgsl_type_name(a.t) + " to a scalar");
            static const std::map<int, std::string> fn = {{Type::F32, "wgb_to_f32"}, {Type::F16, "wgb_to_f16"}, {Type::I32, "wgb_to_i32"}, {Type::U32, "wgb_to_u32"}, {Type::Bool, "wgb_to_bool"}};
            v.s = fn.at(to.k) + "(" + a.s + ")"; v.t = to; return v;
        }
        if (to.k == Type::Vec) {
            // component type: explicit, or inferred from the arguments
            Type elem = *to.elem;
            if (e->ctor_infer) {
                elem = T(Type::AInt);
                bool any = false;
                for (auto& a : args) {
                    const Type& s = a.t.scalar();
                    if (!any) { elem = s; any = true; continue; }
                    if (elem.k == Type::AInt && s.k != Type::AInt) elem = s;
                    else if (elem.k == Type::AFloat && (s.k == Type::F32 || s.k == Type::F16)) elem = s;
                }
                if (!any) elem = T(Type::F32);
                elem = concretize(elem);
            }
            const Type vt = vec_t(to.n, elem);
            v.t = vt;
            if (args.empty()) { v.s = cuda_type(vt, line) + "()"; return v; }
            int comps = 0;
            std::string s;
            const bool splat = args.size() == 1 && args[0].t.is_scalar();
            for (size_t i = 0; i < args.size(); i++) {
                Val a = args[i];
                if (a.t.is_scalar()) { a = coerce(a, elem, line); comps += 1; }
                else if (a.t.k == Type::Vec) {
                    if (args.size() == 1 && a.t.n == to.n && !same(concretize(*a.t.elem), elem)) {
                        // vector conversion, e.g. vec3f(vec3i): component-wise cast
                        a = concrete(a, line);
                        static const std::map<int, std::string> fn = {{Type::F32, "wgb_to_f32"}, {Type::F16, "wgb_to_f16"}, {Type::I32, "wgb_to_i32"}, {Type::U32, "wgb_to_u32"}, {Type::Bool, "wgb_to_bool"}};
                        static const char* comp = "xyzw";
                        std::string c;
                        for (int k = 0; k < to.n; k++) { if (k) c += ", "; c += fn.at(elem.k) + "(wgb_t." + comp[k] + ")"; }
                        v.s = "([&] { const auto wgb_t = " + a.s + "; return " + cuda_type(vt, line) + "(" + c + "); }())";
                        return v;
                    }
                    a = coerce(a, vec_t(a.t.n, elem), line);
                    comps += a.t.n;
                } else err(line, "invalid vector constructor argument " + wgsl_type_name(a.t));
                if (i) s += ", ";
                s += a.s;
            }
            if (!splat && comps != to.n) err(line, "vector constructor has " + std::to_string(comps) + " components, expected " + std::to_string(to.n));
            v.s = cuda_type(vt, line) + "(" + s + ")";
            return v;
        }
        if (to.k == Type::Mat) {
            const Type mt = mat_t(to.n, to.rows, T(Type::F32));
            v.t = mt;
            const std::string name = cuda_type(mt, line);
            if (args.empty()) { v.s = name + "()"; return v; }
            std::string s;
            if ((int)args.size() == to.n && args[0].t.k == Type::Vec) {
                for (size_t i = 0; i < args.size(); i++) { Val a = coerce(args[i], vec_t(to.rows, T(Type::F32)), line); if (i) s += ", "; s += a.s; }
                v.s = name + "(" + s + ")";
                return v;
            }
            if ((int)args.size() == to.n * to.rows) {   // from scalars, column major (compose.rs:145 is a todo!() in the reference)
                for (int c = 0; c < to.n; c++) {
                    if (c) s += ", ";
                    s += "vec" + std::to_string(to.rows) + "f(";
                    for (int r = 0; r < to.rows; r++) { Val a = coerce(args[c * to.rows + r], T(Type::F32), line); if (r) s += ", "; s += a.s; }
                    s += ")";
                }
                v.s = name + "(" + s + ")";
                return v;
            }
            err(line, "unsupported matrix constructor");
        }
        if (to.k == Type::Array) {
            Type elem = *to.elem;
            int count = to.count;
            if (e->ctor_infer) {
                if (args.empty()) err(line, "cannot infer the type of an empty array constructor");
                elem = concretize(args[0].t); count = (int)args.size();
            }
            const Type at = array_t(elem, count);
            v.t = at;
            if (args.empty()) { v.s = cuda_type(at, line) + "()"; return v; }
            if ((int)args.size() != count) err(line, "array constructor needs " + std::to_string(count) + " elements");
            std::string s;
            for (size_t i = 0; i < args.size(); i++) { Val a = coerce(args[i], elem, line); if (i) s += ", "; s += a.s; }
            v.s = cuda_type(at, line) + "{{" + s + "}}";
            return v;
        }
        if (to.k == Type::Struct) {
            v.t = to;
            if (args.empty()) { v.s = to.st->name + "()"; return v; }
            if (args.size() != to.st->members.size()) err(line, "struct constructor needs " + std::to_string(to.st->members.size()) + " values");
            std::string s;
            for (size_t i = 0; i < args.size(); i++) { Val a = coerce(args[i], to.st->members[i].type, line); if (i) s += ", "; s += a.s; }
            v.s = to.st->name + "{" + s + "}";
            return v;
        }
        err(line, "cannot construct " + wgsl_type_name(to));
    }

    Val call(const ExprP& e) {
        const int line = e->line;
        const std::string& name = e->name;
        // user functions
        auto fi = fn_by_name.find(name);
        if (fi != fn_by_name.end()) {
            const Function* f = fi->second;
            if (f->stage()) err(line, "entry points cannot be called");
            if (e->args.size() != f->params.size()) err(line, "wrong number of arguments to '" + name + "'");
            std::string s = name + "(wgb, wgb_inv";
            for (size_t i = 0; i < e->args.size(); i++) { Val a = coerce(expr(e->args[i]), f->params[i].type, line); s += ", " + a.s; }
            Val v; v.s = s + ")"; v.t = f->ret;
            return v;
        }
        if (name == "bitcast") {
            Val a = concrete(expr(e->args[0]), line);
            Val v; v.t = e->ctor;
            if (!(a.t.is_scalar() && v.t.is_scalar())) err(line, "bitcast is supported between 32-bit scalars only");
            v.s = "wgb_bitcast<" + cuda_type(v.t, line) + ">(" + a.s + ")";
            return v;
        }
        std::vector<Val> args;
        auto resource_arg = [&](const ExprP& a, Type::K kind) -> const Global* {
            if (a->k != Expr::Ident) err(line, "expected a resource variable");
            auto gi = global_by_name.find(a->name);
            if (gi == global_by_name.end() || gi->second->type.k != kind) err(line, "'" + a->name + "' is not a " + (kind == Type::Texture2D ? "texture" : "sampler"));
            return gi->second;
        };
        if (name == "textureSample" || name == "textureSampleLevel") {   // expression/image.rs:33-94: 2-D, nearest, mip 0
            if (e->args.size() < 3) err(line, name + " needs (texture, sampler, coords)");
            const Global* tg = resource_arg(e->args[0], Type::Texture2D);
            const Global* sg = resource_arg(e->args[1], Type::Sampler);
            Val uv = coerce(expr(e->args[2]), vec_t(2, T(Type::F32)), line);
            if (name == "textureSample" && e->args.size() > 3) err(line, "textureSample with offset is not supported (binding.rs:112 todo!())");
            Val v; v.t = vec_t(4, T(Type::F32));
            v.s = "wgb_texture_sample(wgb, " + std::to_string(tg->group) + ", " + std::to_string(tg->binding) + ", " + std::to_string(sg->group) + ", " + std::to_string(sg->binding) + ", " + uv.s + ")";
            return v;
        }
        if (name == "textureDimensions") {
            const Global* tg = resource_arg(e->args[0], Type::Texture2D);
            Val v; v.t = vec_t(2, T(Type::U32));
            v.s = "wgb_texture_dimensions(wgb, " + std::to_string(tg->group) + ", " + std::to_string(tg->binding) + ")";
            return v;
        }
        if (name == "textureLoad") {
            const Global* tg = resource_arg(e->args[0], Type::Texture2D);
            Val c = concrete(expr(e->args[1]), line);
            if (!(c.t.k == Type::Vec && c.t.n == 2)) err(line, "textureLoad needs vec2 integer coordinates");
            Val v; v.t = vec_t(4, T(Type::F32));
            v.s = "wgb_texture_load(wgb, " + std::to_string(tg->group) + ", " + std::to_string(tg->binding) + ", " + c.s + ")";
            return v;
        }
        if (name == "arrayLength") {
            // Expression::ArrayLength: elements of the runtime-sized array at the end of a storage binding,
            // from the bound size -- (binding size - offset of the array) / element stride
            if (e->args.size() != 1 || e->args[0]->k != Expr::Unary || e->args[0]->name != "&") err(line, "arrayLength takes a pointer to a runtime-sized array");
            const Global* g = nullptr; uint32_t off = 0; std::string dyn; Type ty;
            if (!resource_chain(e->args[0]->args[0], g, off, dyn, ty) || ty.k != Type::Array || ty.count != 0 || !dyn.empty())
                err(line, "arrayLength takes a pointer to a runtime-sized array of a storage buffer");
            uint32_t a, sz; layout_of(*ty.elem, a, sz, line);
            const uint32_t stride = (sz + a - 1) / a * a;
            Val v; v.t = T(Type::U32);
            v.s = "wgb_array_length(wgb, " + std::to_string(g->group) + ", " + std::to_string(g->binding) + ", " + std::to_string(off) + "u, " + std::to_string(stride) + "u)";
            return v;
        }
        for (auto& a : e->args) args.push_back(expr(a));
        if (name == "select") {
            if (args.size() != 3) err(line, "select needs 3 arguments");
            Val f = args[0], t = args[1], c = args[2];
            Type rt = f.t.is_abstract() ? (t.t.is_abstract() ? concretize((f.t.scalar().k == Type::AFloat || t.t.scalar().k == Type::AFloat) ? (f.t.k == Type::Vec ? vec_t(f.t.n, T(Type::AFloat)) : T(Type::AFloat)) : f.t) : t.t) : f.t;
            f = coerce(f, rt, line); t = coerce(t, rt, line);
            if (c.t.k != Type::Bool && !(c.t.k == Type::Vec && c.t.elem->k == Type::Bool)) err(line, "select condition must be bool");
            Val v; v.t = rt; v.s = "wgb_select(" + f.s + ", " + t.s + ", " + c.s + ")";
            return v;
        }
        // integer bit builtins: same type in and out; offsets / counts are u32
        {
            static const char* bit1[] = {"countOneBits", "countLeadingZeros", "countTrailingZeros", "firstLeadingBit", "firstTrailingBit", "reverseBits"};
            bool is_bit1 = false;
            for (const char* b : bit1) is_bit1 = is_bit1 || name == b;
            if (is_bit1 || name == "extractBits" || name == "insertBits") {
                const size_t want_n = is_bit1 ? 1 : name == "extractBits" ? 3 : 4;
                if (args.size() != want_n) err(line, name + " takes " + std::to_string(want_n) + " argument(s)");
                Val e0 = concrete(args[0], line);
                const Type::K sk = e0.t.scalar().k;
                if (sk != Type::I32 && sk != Type::U32) err(line, name + " needs integer arguments");
                std::string s = "wgb_" + name + "(" + e0.s;
                size_t k = 1;
                if (name == "insertBits") { s += ", " + coerce(args[1], e0.t, line).s; k = 2; }
                for (; k < args.size(); k++) s += ", " + coerce(args[k], T(Type::U32), line).s;
                Val v; v.t = e0.t; v.s = s + ")";
                return v;
            }
        }
        // math builtins (all todo!() in the reference, expression/math.rs:23,29)
        struct B { const char* name; int nargs; int ret; };   // ret: 0 same as arg0, 1 scalar of arg0, 2 bool, 3 vec3
        static const B table[] = {
            {"abs", 1, 0}, {"min", 2, 0}, {"max", 2, 0}, {"clamp", 3, 0}, {"saturate", 1, 0}, {"floor", 1, 0}, {"ceil", 1, 0}, {"round", 1, 0},
            {"trunc", 1, 0}, {"fract", 1, 0}, {"sqrt", 1, 0}, {"inverseSqrt", 1, 0}, {"sin", 1, 0}, {"cos", 1, 0}, {"tan", 1, 0}, {"asin", 1, 0},
            {"acos", 1, 0}, {"atan", 1, 0}, {"atan2", 2, 0}, {"sinh", 1, 0}, {"cosh", 1, 0}, {"tanh", 1, 0}, {"exp", 1, 0}, {"exp2", 1, 0},
            {"log", 1, 0}, {"log2", 1, 0}, {"pow", 2, 0}, {"sign", 1, 0}, {"step", 2, 0}, {"smoothstep", 3, 0}, {"mix", 3, 0}, {"fma", 3, 0},
            {"degrees", 1, 0}, {"radians", 1, 0}, {"dot", 2, 1}, {"length", 1, 1}, {"distance", 2, 1}, {"normalize", 1, 0}, {"cross", 2, 0},
            {"reflect", 2, 0}, {"all", 1, 2}, {"any", 1, 2}, {"transpose", 1, 0}, {nullptr, 0, 0}};
        for (int i = 0; table[i].name; i++) {
            if (name != table[i].name) continue;
            if ((int)args.size() != table[i].nargs) err(line, name + " takes " + std::to_string(table[i].nargs) + " argument(s)");
            // unify argument types: the widest non-abstract argument decides
            Type want = args[0].t;
            for (auto& a : args) { if (want.is_abstract() && !a.t.is_abstract()) want = a.t; if (a.t.k == Type::Vec && want.is_scalar()) want = vec_t(a.t.n, want.scalar()); }
            want = concretize(want);
            const bool int_ok = name == "abs" || name == "min" || name == "max" || name == "clamp" || name == "sign" || name == "dot";
            if (want.scalar().k == Type::F16 && !(name == "abs" || name == "min" || name == "max" || name == "clamp"))
                err(line, name + " is not supported for f16 arguments (abs, min, max, clamp and select are)");
            if (table[i].ret != 2 && want.scalar().k != Type::F32 && want.scalar().k != Type::F16 && !int_ok) {
                if (want.scalar().k == Type::I32 && args[0].t.is_abstract()) want = want.k == Type::Vec ? vec_t(want.n, T(Type::F32)) : T(Type::F32);
                else err(line, name + " needs float arguments");
            }
            std::string s = "wgb_" + name + "(";
            for (size_t k = 0; k < args.size(); k++) {
                Val a = args[k];
                // scalar third argument of mix / scalar edges are allowed to stay scalar
                Type target = want;
                if (a.t.is_scalar() && want.k == Type::Vec && (name == "mix" || name == "clamp" || name == "min" || name == "max" || name == "step" || name == "smoothstep" || name == "pow"))
                    target = want.scalar();
                a = coerce(a, target, line);
                if (k) s += ", ";
                s += a.s;
            }
            Val v; v.s = s + ")";
            v.t = table[i].ret == 0 ? want : table[i].ret == 1 ? want.scalar() : T(Type::Bool);
            return v;
        }
        err(line, "unknown function '" + name + "'");
    }

    // ---- statements ----
    std::string ind(int d) const { return std::string(d * 4, ' '); }
    std::string zero_return() const {
        if (cur_fn->ret.k == Type::Void) return "return;";
        return "return " + cuda_type(cur_fn->ret, cur_fn->line) + "();";
    }
    bool expr_calls_discarding(const ExprP& e) const {
        if (!e) return false;
        if (e->k == Expr::Call) { auto f = fn_by_name.find(e->name); if (f != fn_by_name.end() && f->second->may_discard) return true; }
        for (auto& a : e->args) if (expr_calls_discarding(a)) return true;
        return false;
    }
    void after_call_check(const ExprP& e, int d) {
        if (expr_calls_discarding(e)) out += ind(d) + "if (wgb_inv.killed) " + zero_return() + "\n";
    }

    void block(const std::vector<StmtP>& b, int d) {
        push();
        for (auto& s : b) stmt(s, d);
        pop();
    }
    std::string simple_stmt_text(const StmtP& s) {   // for-header statements, no trailing ';'
        std::string save = out;
        out.clear();
        stmt(s, 0);
        std::string t = out;
        out = save;
        while (!t.empty() && (t.back() == '\n' || t.back() == ';' || t.back() == ' ')) t.pop_back();
        return t;
    }
    void stmt(const StmtP& s, int d) {
        const int line = s->line;
        switch (s->k) {
            case Stmt::Block: out += ind(d) + "{\n"; block(s->body, d + 1); out += ind(d) + "}\n"; return;
            case Stmt::Let: case Stmt::Const: case Stmt::Var: {
                Type ty;
                std::string init;
                if (s->a) {
                    Val v = expr(s->a);
                    if (s->has_type) v = coerce(v, s->decl_type, line); else v = concrete(v, line);
                    ty = v.t; init = v.s;
                } else { ty = s->decl_type; init = cuda_type(ty, line) + "()"; }
                if (ty.k == Type::Void) err(line, "cannot declare a variable of type void");
                out += ind(d) + (s->k == Stmt::Var ? "" : "const ") + cuda_type(ty, line) + " " + s->name + " = " + init + ";\n";
                declare(s->name, ty, s->k != Stmt::Var, line);
                after_call_check(s->a, d);
                return;
            }
            case Stmt::Assign: {
                Val rhs = expr(s->b);
                // the left-hand side must be a local / private variable access chain
                Val lhs = lvalue(s->a);
                std::string op = s->op;
                if (op == "=") { rhs = coerce(rhs, lhs.t, line); out += ind(d) + lhs.s + " = " + rhs.s + ";\n"; }
                else {
                    // a op= b  ==>  a = a op b with the full operator semantics
                    auto bin = std::make_shared<Expr>(); bin->k = Expr::Binary; bin->line = line; bin->name = op.substr(0, op.size() - 1); bin->args = {s->a, s->b};
                    Val r = coerce(expr(bin), lhs.t, line);
                    out += ind(d) + lhs.s + " = " + r.s + ";\n";
                }
                after_call_check(s->b, d);
                return;
            }
            case Stmt::Incr: {
                Val lhs = lvalue(s->a);
                if (!(lhs.t.k == Type::I32 || lhs.t.k == Type::U32)) err(line, "'" + s->op + "' needs an integer variable");
                out += ind(d) + lhs.s + " = (" + lhs.s + (s->op == "++" ? " + 1" : " - 1") + (lhs.t.k == Type::U32 ? "u" : "") + ");\n";
                return;
            }
            case Stmt::If: {
                Val c = expr(s->a);
                if (c.t.k != Type::Bool) err(line, "if condition must be bool");
                out += ind(d) + "if (" + c.s + ") {\n";
                block(s->body, d + 1);
                if (!s->else_body.empty()) { out += ind(d) + "} else {\n"; block(s->else_body, d + 1); }
                out += ind(d) + "}\n";
                return;
            }
            case Stmt::For: {
                push();
                out += ind(d) + "{\n";
                if (s->init) stmt(s->init, d + 1);
                std::string cond = "true";
                if (s->a) { Val c = expr(s->a); if (c.t.k != Type::Bool) err(line, "for condition must be bool"); cond = c.s; }
                std::string upd;
                if (s->update) upd = simple_stmt_text(s->update);
                out += ind(d + 1) + "for (; " + cond + "; " + upd + ") {\n";
                continue_labels.push_back("");
                block(s->body, d + 2);
                continue_labels.pop_back();
                out += ind(d + 1) + "}\n" + ind(d) + "}\n";
                pop();
                return;
            }
            case Stmt::While: {
                Val c = expr(s->a);
                if (c.t.k != Type::Bool) err(line, "while condition must be bool");
                out += ind(d) + "while (" + c.s + ") {\n";
                continue_labels.push_back("");
                block(s->body, d + 1);
                continue_labels.pop_back();
                out += ind(d) + "}\n";
                return;
            }
            case Stmt::Loop: {
                const std::string label = "wgb_continuing_" + std::to_string(label_counter++);
                out += ind(d) + "for (;;) {\n";
                push();
                continue_labels.push_back(s->continuing.empty() ? "" : label);
                for (auto& c : s->body) stmt(c, d + 1);
                continue_labels.pop_back();
                if (!s->continuing.empty()) {
                    out += ind(d + 1) + label + ":;\n";
                    continue_labels.push_back("!");   // continue is not allowed inside continuing
                    for (auto& c : s->continuing) stmt(c, d + 1);
                    continue_labels.pop_back();
                }
                pop();
                out += ind(d) + "}\n";
                return;
            }
            case Stmt::BreakIf: {
                Val c = expr(s->a);
                if (c.t.k != Type::Bool) err(line, "break if condition must be bool");
                out += ind(d) + "if (" + c.s + ") break;\n";
                return;
            }
            case Stmt::Switch: {
                Val sel = concrete(expr(s->a), line);
                if (!(sel.t.k == Type::I32 || sel.t.k == Type::U32)) err(line, "switch selector must be an integer");
                out += ind(d) + "switch (" + sel.s + ") {\n";
                bool has_default = false;
                continue_labels.push_back(continue_labels.empty() ? "" : continue_labels.back());
                for (auto& c : s->cases) {
                    for (auto& se : c.sel) { Val cv = coerce(expr(se), sel.t, line); out += ind(d + 1) + "case " + cv.s + ":\n"; }
                    if (c.is_default) { out += ind(d + 1) + "default:\n"; has_default = true; }
                    out += ind(d + 1) + "{\n";
                    block(c.body, d + 2);
                    out += ind(d + 2) + "break;\n" + ind(d + 1) + "}\n";
                }
                continue_labels.pop_back();
                if (!has_default) out += ind(d + 1) + "default: break;\n";
                out += ind(d) + "}\n";
                return;
            }
            case Stmt::Break: out += ind(d) + "break;\n"; return;
            case Stmt::Continue:
                if (!continue_labels.empty() && continue_labels.back() == "!") err(line, "continue inside a continuing block");
                if (!continue_labels.empty() && !continue_labels.back().empty()) out += ind(d) + "goto " + continue_labels.back() + ";\n";
                else out += ind(d) + "continue;\n";
                return;
            case Stmt::Discard:
                if (stage != 2) err(line, "discard outside a fragment entry point's call graph");
                out += ind(d) + "{ wgb_inv.killed = true; " + zero_return() + " }\n";   // statement/kill.rs:22-36
                return;
            case Stmt::Return: {
                if (!s->a) { if (cur_fn->ret.k != Type::Void) err(line, "missing return value"); out += ind(d) + "return;\n"; return; }
                Val v = coerce(expr(s->a), cur_fn->ret, line);
                if (expr_calls_discarding(s->a)) {
                    out += ind(d) + "{ const " + cuda_type(cur_fn->ret, line) + " wgb_r = " + v.s + "; if (wgb_inv.killed) " + zero_return() + " return wgb_r; }\n";
                } else out += ind(d) + "return " + v.s + ";\n";
                return;
            }
            case Stmt::CallS: {
                Val v = expr(s->a);
                out += ind(d) + "(void)(" + v.s + ");\n";
                after_call_check(s->a, d);
                return;
            }
        }
    }
    Val lvalue(const ExprP& e) {
        const int line = e->line;
        if (e->k == Expr::Ident) {
            if (const Var* var = lookup(e->name)) { if (var->is_const) err(line, "cannot assign to '" + e->name + "'"); Val v; v.s = var->cname; v.t = var->type; return v; }
            auto gi = global_by_name.find(e->name);
            if (gi != global_by_name.end() && gi->second->space == "private") { Val v; v.s = "wgb_inv." + e->name; v.t = gi->second->type; return v; }
            if (gi != global_by_name.end()) err(line, "buffers are read-only in shaders (runtime.rs:1528)");
            err(line, "unknown identifier '" + e->name + "'");
        }
        if (e->k == Expr::Member) {
            Val a = lvalue(e->args[0]);
            if (a.t.k == Type::Struct) {
                for (auto& mb : a.t.st->members) if (mb.name == e->name) { Val v; v.s = a.s + "." + mb.name; v.t = mb.type; return v; }
                err(line, "struct " + a.t.st->name + " has no member '" + e->name + "'");
            }
            if (a.t.k == Type::Vec) {
                std::vector<int> idx;
                if (!is_swizzle(e->name, a.t.n, idx) || idx.size() != 1) err(line, "cannot assign to swizzle '." + e->name + "'");
                Val v; v.s = a.s + "." + "xyzw"[idx[0]]; v.t = *a.t.elem; return v;
            }
            err(line, "cannot assign to member of " + wgsl_type_name(a.t));
        }
        if (e->k == Expr::Index) {
            Val a = lvalue(e->args[0]);
            Val i = expr(e->args[1]);
            if (!i.is_const_num) i = concrete(i, line);
            Val v;
            int cnt;
            if (a.t.k == Type::Vec) { v.t = *a.t.elem; cnt = a.t.n; }
            else if (a.t.k == Type::Mat) { v.t = vec_t(a.t.rows, *a.t.elem); cnt = a.t.n; }
            else if (a.t.k == Type::Array) { v.t = *a.t.elem; cnt = a.t.count; }
            else err(line, "cannot index " + wgsl_type_name(a.t));
            if (i.is_const_num) v.s = a.s + "[" + std::to_string((int)i.num) + "]";
            else v.s = a.s + "[min((u32)(" + i.s + "), " + std::to_string(cnt - 1) + "u)]";
            return v;
        }
        err(line, "expression is not assignable");
    }

    // ---- call graph: which functions may discard ----
    bool stmt_discards(const std::vector<StmtP>& b, const std::set<std::string>& discarding) const {
        for (auto& s : b) {
            if (!s) continue;
            if (s->k == Stmt::Discard) return true;
            auto ed = [&](const ExprP& e) { return e && expr_calls(e, discarding); };
            if (ed(s->a) || ed(s->b)) return true;
            if (s->init && stmt_discards({s->init}, discarding)) return true;
            if (s->update && stmt_discards({s->update}, discarding)) return true;
            if (stmt_discards(s->body, discarding) || stmt_discards(s->else_body, discarding) || stmt_discards(s->continuing, discarding)) return true;
            for (auto& c : s->cases) if (stmt_discards(c.body, discarding)) return true;
        }
        return false;
    }
    // does the block call one of `names`? (the walk of stmt_discards without the discard statement itself)
    bool stmt_calls(const std::vector<StmtP>& b, const std::set<std::string>& names) const {
        for (auto& s : b) {
            if (!s) continue;
            auto ec = [&](const ExprP& e) { return e && expr_calls(e, names); };
            if (ec(s->a) || ec(s->b)) return true;
            if (s->init && stmt_calls({s->init}, names)) return true;
            if (s->update && stmt_calls({s->update}, names)) return true;
            if (stmt_calls(s->body, names) || stmt_calls(s->else_body, names) || stmt_calls(s->continuing, names)) return true;
            for (auto& c : s->cases) if (stmt_calls(c.body, names)) return true;
        }
        return false;
    }
    bool expr_calls(const ExprP& e, const std::set<std::string>& names) const {
        if (e->k == Expr::Call && names.count(e->name)) return true;
        for (auto& a : e->args) if (a && expr_calls(a, names)) return true;
        return false;
    }

    // ---- module emission ----
    static const Attr* find_attr(const std::vector<Attr>& as, const char* n) { for (auto& a : as) if (a.name == n) return &a; return nullptr; }

    std::string run() {
        for (auto& sd : m.structs) layout_struct(*sd, 0);
        for (auto& g : m.globals) global_by_name[g.name] = &g;
        for (auto& f : m.functions) fn_by_name[f->name] = f.get();
        const Function* ep = nullptr;
        for (auto& f : m.functions) if (f->name == entry && f->stage() == stage) ep = f.get();
        if (!ep) {
            for (auto& f : m.functions) if (f->name == entry) err(f->line, "entry point '" + entry + "' is not a " + (stage == 1 ? "vertex" : "fragment") + " entry point");
            throw std::runtime_error("WGSL: no entry point named '" + entry + "'");
        }
        // may_discard fixed point
        std::set<std::string> discarding;
        for (bool changed = true; changed;) {
            changed = false;
            for (auto& f : m.functions)
                if (!discarding.count(f->name) && stmt_discards(f->body, discarding)) { discarding.insert(f->name); f->may_discard = true; changed = true; }
        }
        for (auto& g : m.globals) if (g.space == "private") has_private = true;
        // only the entry point's call graph is emitted: a module shared by both stages may hold helpers that are valid in
        // one of them only (a function that discards must not stop the vertex stage of the same module from translating)
        std::set<std::string> reachable = {ep->name};
        for (bool changed = true; changed;) {
            changed = false;
            for (auto& f : m.functions) {
                if (reachable.count(f->name) || f->stage()) continue;
                for (auto& g : m.functions)
                    if (reachable.count(g->name) && stmt_calls(g->body, {f->name})) { reachable.insert(f->name); changed = true; break; }
            }
        }

        out += "// wgsl2cuda: stage=" + std::string(stage == 1 ? "vertex" : "fragment") + " entry=" + entry + "\n";
        out += "namespace " + ns + " {\n";
        scopes.clear();
        push();
        // per-invocation state: the discard flag and module-scope private variables
        std::string inv = "struct WgbInvocation {\n    bool killed = false;\n";
        for (auto& item : m.decl_order) {
            const std::string kind = item.substr(0, 2), name = item.substr(2);
            if (kind == "s:") {
                for (auto& sd : m.structs) if (sd->name == name) {
                    out += "struct " + sd->name + " {";
                    for (auto& mb : sd->members) {
                        if (mb.type.k == Type::Array && mb.type.count == 0) continue;   // a runtime-sized tail lives in the buffer only (reached through access chains)
                        out += " " + cuda_type(mb.type, mb.line) + " " + mb.name + ";";
                    }
                    out += " };\n";
                }
            } else if (kind == "g:") {
                Global& g = const_cast<Global&>(*global_by_name[name]);
                if (g.space == "const") {
                    Val v = expr(g.init);
                    if (g.type.k != Type::Void) v = coerce(v, g.type, g.line); else v = concrete(v, g.line);
                    g.type = v.t;
                    if (g.type.is_scalar()) out += "static constexpr " + cuda_type(g.type, g.line) + " " + g.name + " = " + v.s + ";\n";
                    else out += "WGB_DEV " + cuda_type(g.type, g.line) + " " + g.name + "() { return " + v.s + "; }\n";
                } else if (g.space == "private") {
                    std::string init = cuda_type(g.type, g.line) + "()";
                    if (g.init) { Val v = coerce(expr(g.init), g.type, g.line); init = v.s; }
                    inv += "    " + cuda_type(g.type, g.line) + " " + g.name + " = " + init + ";\n";
                } else if (g.space == "uniform" || g.space == "storage") {
                    if (g.group < 0 || g.binding < 0) err(g.line, "resource '" + g.name + "' needs @group and @binding");
                    if (g.group >= 4 || g.binding >= 4) err(g.line, "resource '" + g.name + "': group/binding indices above 3 are not supported");
                    uint32_t a, sz; layout_of(g.type, a, sz, g.line);   // validates host-shareability
                } else {
                    if (g.type.k != Type::Texture2D && g.type.k != Type::Sampler) err(g.line, "module-scope var '" + g.name + "' needs an address space");
                    if (g.group < 0 || g.binding < 0 || g.group >= 4 || g.binding >= 4) err(g.line, "resource '" + g.name + "' needs @group/@binding below 4");
                }
            }
        }
        inv += "};\n";
        // functions, in declaration order (WGSL allows use before declaration: emit prototypes first)
        std::string protos, bodies;
        for (auto& f : m.functions) {
            if (f->stage() && f.get() != ep) continue;
            if (!reachable.count(f->name)) continue;
            cur_fn = f.get();
            std::string sig = "WGB_DEV " + cuda_type(f->ret, f->line) + " " + f->name + "(const WgbDraw& wgb, WgbInvocation& wgb_inv";
            push();
            for (auto& pr : f->params) { sig += ", " + cuda_type(pr.type, pr.line) + " " + pr.name; declare(pr.name, pr.type, true, pr.line); }
            sig += ")";
            protos += sig + ";\n";
            const std::string save = out;
            out.clear();
            label_counter = 0;
            block(f->body, 1);
            const std::string body = out;
            out = save;
            pop();
            const bool ends_with_return = !f->body.empty() && f->body.back() && f->body.back()->k == Stmt::Return;
            bodies += sig + " {\n" + body + ((f->ret.k == Type::Void || ends_with_return) ? "" : "    " + zero_return() + "\n") + "}\n";
        }
        out += inv + protos + bodies;
        out += "}  // namespace " + ns + "\n";
        pop();
        if (stage == 1) emit_vertex_glue(*ep); else emit_fragment_glue(*ep);
        return out;
    }

    // inter-stage slots: locations packed in declaration order, each aligned like naga's layouter
    // (naga-cranelift/src/bindings.rs:307-346), in units of 4 bytes
    static int slots_of(const Type& t, int line) {
        if (t.is_scalar()) return 1;
        if (t.k == Type::Vec) return t.n;
        err(line, "inter-stage variables must be scalars or vectors");
    }
    static int slot_align(const Type& t) { return t.k == Type::Vec ? (t.n == 2 ? 2 : 4) : 1; }

    struct IoItem { std::string access; Type type; std::string builtin; int location = -1; std::string interp, sampling; int line = 0; };
    void collect_io(const std::string& prefix, const Type& t, const std::vector<Attr>& attrs, int line, std::vector<IoItem>& items) {
        if (t.k == Type::Struct) {
            for (auto& mb : t.st->members) collect_io(prefix + "." + mb.name, mb.type, mb.attrs, mb.line, items);
            return;
        }
        if (t.scalar().k == Type::F16) err(line, "f16 entry point inputs / outputs are not supported (convert to f32 at the stage boundary)");
        IoItem it; it.access = prefix; it.type = t; it.line = line;
        if (const Attr* b = find_attr(attrs, "builtin")) { if (b->args.empty()) err(line, "@builtin needs a name"); it.builtin = b->args[0]; }
        if (const Attr* l = find_attr(attrs, "location")) { if (l->args.empty()) err(line, "@location needs an index"); it.location = atoi(l->args[0].c_str()); }
        if (const Attr* i = find_attr(attrs, "interpolate")) { if (!i->args.empty()) it.interp = i->args[0]; if (i->args.size() > 1) it.sampling = i->args[1]; }
        if (it.builtin.empty() && it.location < 0) err(line, "entry point input/output '" + prefix + "' needs @builtin or @location");
        items.push_back(it);
    }

    void emit_vertex_glue(const Function& f) {
        std::vector<IoItem> outs;
        collect_io("r", f.ret, f.ret_attrs, f.line, outs);
        int slot = 0;
        std::string defs, puts;
        bool have_position = false;
        for (auto& o : outs) {
            if (o.builtin == "position") { puts += "    position = " + o.access + ";\n"; have_position = true; continue; }
            if (!o.builtin.empty()) err(o.line, "vertex output builtin '" + o.builtin + "' is not supported");
            const int n = slots_of(o.type, o.line), al = slot_align(o.type);
            slot = (slot + al - 1) / al * al;
            defs += "#define WGB_VS_LOC" + std::to_string(o.location) + "_SLOT " + std::to_string(slot) + "\n";
            puts += "    wgb_put(vary, WGB_VS_LOC" + std::to_string(o.location) + "_SLOT, " + o.access + ");\n";
            slot += n;
        }
        if (!have_position) err(f.line, "vertex entry point must return @builtin(position)");
        out += "#define WGB_VS_VARYING_SLOTS " + std::to_string(slot) + "\n" + defs;
        out += "WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {\n";
        out += "    " + ns + "::WgbInvocation wgb_inv;\n";
        std::string call = ns + "::" + f.name + "(wgb, wgb_inv";
        for (size_t i = 0; i < f.params.size(); i++) {
            const Param& pr = f.params[i];
            const std::string an = "a" + std::to_string(i);
            out += "    " + (pr.type.k == Type::Struct ? ns + "::" : std::string()) + cuda_type(pr.type, pr.line) + " " + an + ";\n";
            std::vector<IoItem> ins;
            collect_io(an, pr.type, pr.attrs, pr.line, ins);
            for (auto& in : ins) {
                if (in.builtin == "vertex_index") out += "    " + in.access + " = vertex_index;\n";
                else if (in.builtin == "instance_index") out += "    " + in.access + " = instance_index;\n";
                else if (!in.builtin.empty()) err(in.line, "vertex input builtin '" + in.builtin + "' is not supported");
                else out += "    " + in.access + " = WGB_FETCH(" + cuda_type(in.type, in.line) + ", " + std::to_string(in.location) + ");\n";
            }
            call += ", " + an;
        }
        out += "    const " + (f.ret.k == Type::Struct ? ns + "::" : std::string()) + cuda_type(f.ret, f.line) + " r = " + call + ");\n";
        out += puts + "}\n";
    }

    void emit_fragment_glue(const Function& f) {
        // inputs
        std::string body, interp = "    return ";
        bool uses_front_facing = false;
        std::string call = ns + "::" + f.name + "(wgb, wgb_inv";
        for (size_t i = 0; i < f.params.size(); i++) {
            const Param& pr = f.params[i];
            const std::string an = "a" + std::to_string(i);
            body += "    " + (pr.type.k == Type::Struct ? ns + "::" : std::string()) + cuda_type(pr.type, pr.line) + " " + an + ";\n";
            std::vector<IoItem> ins;
            collect_io(an, pr.type, pr.attrs, pr.line, ins);
            for (auto& in : ins) {
                if (in.builtin == "position") body += "    " + in.access + " = fi.position;\n";
                else if (in.builtin == "front_facing") { body += "    " + in.access + " = fi.front_facing;\n"; uses_front_facing = true; }
                else if (in.builtin == "primitive_index") body += "    " + in.access + " = fi.primitive_index;\n";
                else if (in.builtin == "sample_index") body += "    " + in.access + " = fi.sample_index;\n";
                else if (in.builtin == "sample_mask") body += "    " + in.access + " = fi.sample_mask;\n";
                else if (!in.builtin.empty()) err(in.line, "fragment input builtin '" + in.builtin + "' is not supported");
                else {
                    const std::string loc = "WGB_VS_LOC" + std::to_string(in.location) + "_SLOT";
                    body += "    " + in.access + " = wgb_get<" + cuda_type(in.type, in.line) + ">(vary, " + loc + ");\n";
                    // Interpolation::from_naga (fragment.rs:282-317): default is perspective
                    int mode = 2;
                    if (in.interp == "flat") mode = 0; else if (in.interp == "linear") mode = 1; else if (in.interp == "perspective" || in.interp.empty()) mode = 2;
                    else err(in.line, "unknown interpolation '" + in.interp + "'");
                    if (in.type.scalar().k != Type::F32 && mode != 0) err(in.line, "Integer types must use flat interpolation");   // fragment.rs:335
                    interp += "(slot >= " + loc + " && slot < " + loc + " + " + std::to_string(slots_of(in.type, in.line)) + ") ? " + std::to_string(mode) + " : ";
                }
            }
            call += ", " + an;
        }
        interp += "0;\n";
        // outputs, visited in declaration order (fragment.rs:457-488)
        std::vector<IoItem> outs;
        if (f.ret.k != Type::Void) collect_io("r", f.ret, f.ret_attrs, f.line, outs);
        int mask = 0;
        bool writes_depth = false, seen_location = false;
        std::string stores;
        for (auto& o : outs) {
            if (o.builtin == "frag_depth") {
                // FragDepth only counts if it precedes the first @location output: the late depth test runs there
                if (!seen_location) { writes_depth = true; stores += "    out.frag_depth = " + o.access + ";\n"; }
                continue;
            }
            if (o.builtin == "sample_mask") continue;
            if (!o.builtin.empty()) err(o.line, "fragment output builtin '" + o.builtin + "' is not supported");
            if (o.location >= 4) err(o.line, "at most 4 colour targets are supported");
            if (!(o.type.k == Type::Vec && o.type.n == 4 && o.type.elem->k == Type::F32)) err(o.line, "colour outputs must be vec4<f32> (fragment.rs:481)");
            seen_location = true;
            mask |= 1 << o.location;
            stores += "    out.color[" + std::to_string(o.location) + "] = " + o.access + ";\n";
        }
        int early = 0;
        if (const Attr* a = find_attr(f.attrs, "early_depth_test")) early = (!a->args.empty() && a->args[0] == "force") ? 1 : 2;
        out += "#define WGB_FS_COLOR_MASK " + std::to_string(mask) + "\n";
        out += "#define WGB_FS_WRITES_FRAG_DEPTH " + std::to_string(writes_depth ? 1 : 0) + "\n";
        out += "#define WGB_FS_MAY_DISCARD " + std::to_string(f.may_discard ? 1 : 0) + "\n";
        out += "#define WGB_FS_EARLY_DEPTH " + std::to_string(early) + "\n";
        out += "#define WGB_FS_USES_FRONT_FACING " + std::to_string(uses_front_facing ? 1 : 0) + "\n";
        out += "WGB_DEV constexpr int wgb_fs_interp(int slot) {\n" + interp + "}\n";
        out += "WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {\n";
        out += "    " + ns + "::WgbInvocation wgb_inv;\n" + body;
        if (f.ret.k == Type::Void) out += "    " + call + ");\n";
        else out += "    const " + (f.ret.k == Type::Struct ? ns + "::" : std::string()) + cuda_type(f.ret, f.line) + " r = " + call + ");\n";
        out += "    if (wgb_inv.killed) return false;\n" + stores + "    return true;\n}\n";
    }
};

}  // namespace

std::string wgb_emit_wgsl(const std::string& wgsl, uint32_t stage, const std::string& entry_point) {
    if (stage != 1 && stage != 2) throw std::runtime_error("WGSL: only vertex and fragment entry points are supported (the reference has no compute stage, device.rs:143-148)");
    Parser ps;
    ps.t = lex(wgsl);
    ps.parse_module();
    Emitter em(ps.m, (int)stage, entry_point);
    return em.run();
}
