// dsl_parse.cpp — lexer, expression parser and the two surface-form parsers.
// Grammar references: pharmsol-dsl/src/lexer.rs:10-62, parser.rs:1022-1261 (expressions and
// precedence), parser.rs:300-1020 (canonical blocks), authoring.rs:362-900 (shorthand lines).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <set>
#include <sstream>

#include "dsl.hpp"

namespace pharmsol {
namespace dsl {

namespace {

enum class Tk { End, Ident, Number, Punct, Newline };
struct Token {
    Tk kind = Tk::End;
    std::string text;
    double num = 0.0;
    int pos = 0;
};

std::vector<Token> lex(const std::string& s, int base, bool keep_newlines) {
    std::vector<Token> out;
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
        const char c = s[i];
        if (c == '\n') {
            if (keep_newlines) out.push_back(Token{Tk::Newline, "\n", 0.0, base + (int)i});
            ++i; continue;
        }
        if (std::isspace((unsigned char)c)) { ++i; continue; }
        if (c == '#') { while (i < n && s[i] != '\n') ++i; continue; }
        if (c == '/' && i + 1 < n && s[i + 1] == '/') { while (i < n && s[i] != '\n') ++i; continue; }
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t j = i;
            while (j < n && (std::isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
            out.push_back(Token{Tk::Ident, s.substr(i, j - i), 0.0, base + (int)i});
            i = j; continue;
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < n && std::isdigit((unsigned char)s[i + 1]))) {
            size_t j = i;
            while (j < n && std::isdigit((unsigned char)s[j])) ++j;
            // a '.' followed by another '.' is the range operator, not a decimal point
            if (j < n && s[j] == '.' && !(j + 1 < n && s[j + 1] == '.')) {
                ++j;
                while (j < n && std::isdigit((unsigned char)s[j])) ++j;
            }
            if (j < n && (s[j] == 'e' || s[j] == 'E')) {
                size_t k = j + 1;
                if (k < n && (s[k] == '+' || s[k] == '-')) ++k;
                if (k < n && std::isdigit((unsigned char)s[k])) {
                    while (k < n && std::isdigit((unsigned char)s[k])) ++k;
                    j = k;
                }
            }
            Token t{Tk::Number, s.substr(i, j - i), 0.0, base + (int)i};
            t.num = std::strtod(t.text.c_str(), nullptr);
            out.push_back(t);
            i = j; continue;
        }
        static const char* two[] = {"->", "==", "!=", "<=", ">=", "&&", "||", ".."};
        bool matched = false;
        for (const char* op : two) {
            if (i + 1 < n && s[i] == op[0] && s[i + 1] == op[1]) {
                out.push_back(Token{Tk::Punct, op, 0.0, base + (int)i});
                i += 2; matched = true; break;
            }
        }
        if (matched) continue;
        if (std::string("+-*/^()[]{},=<>!~@:;").find(c) != std::string::npos) {
            out.push_back(Token{Tk::Punct, std::string(1, c), 0.0, base + (int)i});
            ++i; continue;
        }
        throw DslError(std::string("unexpected character `") + c + "`", base + (int)i);
    }
    out.push_back(Token{Tk::End, "", 0.0, base + (int)n});
    return out;
}

ExprP mk(Expr::Kind k, int pos) { auto e = std::make_shared<Expr>(); e->kind = k; e->pos = pos; return e; }

struct Parser {
    std::vector<Token> toks;
    size_t i = 0;
    explicit Parser(std::vector<Token> t) : toks(std::move(t)) {}
    const Token& peek(size_t k = 0) const { return toks[std::min(i + k, toks.size() - 1)]; }
    Token bump() { Token t = peek(); if (i < toks.size() - 1) ++i; return t; }
    bool at_punct(const char* p) const { return peek().kind == Tk::Punct && peek().text == p; }
    bool at_ident(const char* p) const { return peek().kind == Tk::Ident && peek().text == p; }
    bool take_punct(const char* p) { if (at_punct(p)) { bump(); return true; } return false; }
    void expect_punct(const char* p) {
        if (!take_punct(p)) throw DslError(std::string("expected `") + p + "`, found `" + peek().text + "`", peek().pos);
    }
    std::string expect_ident() {
        if (peek().kind != Tk::Ident) throw DslError("expected identifier, found `" + peek().text + "`", peek().pos);
        return bump().text;
    }
    void skip_newlines() { while (peek().kind == Tk::Newline || at_punct(",") || at_punct(";")) bump(); }
    void skip_nl_only() { while (peek().kind == Tk::Newline) bump(); }

    // binary precedence (parser.rs:1242-1261): || 1, && 2, == != 3, < <= > >= 4, + - 5, * / 6, ^ 7 (right assoc)
    static int prec(const std::string& op) {
        if (op == "||") return 1;
        if (op == "&&") return 2;
        if (op == "==" || op == "!=") return 3;
        if (op == "<" || op == "<=" || op == ">" || op == ">=") return 4;
        if (op == "+" || op == "-") return 5;
        if (op == "*" || op == "/") return 6;
        if (op == "^") return 7;
        return 0;
    }
    // Recursion guard: a shared library must answer degenerate input ("((((...", 10^5 chained operators) with an error,
    // not with a stack overflow of the host process.  Every recursive production passes through here; chained
    // left-associative operators are bounded by the number of nodes on the spine (the emitter recurses over it).
    static constexpr int kMaxDepth = 400;
    int depth = 0;
    struct DepthGuard {
        Parser& p;
        explicit DepthGuard(Parser& pp) : p(pp) { if (++p.depth > kMaxDepth) throw DslError("expression or block nesting is too deep (limit " + std::to_string(kMaxDepth) + ")", p.peek().pos); }
        ~DepthGuard() { --p.depth; }
    };
    ExprP parse_expr(int min_prec = 1) {
        DepthGuard guard(*this);
        ExprP lhs = parse_unary();
        int spine = 0;
        while (peek().kind == Tk::Punct) {
            if (++spine > 4 * kMaxDepth) throw DslError("expression has too many chained operators (limit " + std::to_string(4 * kMaxDepth) + ")", peek().pos);
            const std::string op = peek().text;
            const int p = prec(op);
            if (p == 0 || p < min_prec) break;
            const int pos = bump().pos;
            ExprP rhs = parse_expr(op == "^" ? p : p + 1);
            ExprP b = mk(Expr::Binary, pos);
            b->name = op; b->args = {lhs, rhs};
            lhs = b;
        }
        return lhs;
    }
    // unary + - ! bind tighter than any binary operator (so -a^2 == (-a)^2)
    ExprP parse_unary() {
        DepthGuard guard(*this);
        if (peek().kind == Tk::Punct && (peek().text == "-" || peek().text == "+" || peek().text == "!")) {
            Token t = bump();
            ExprP u = mk(Expr::Unary, t.pos);
            u->name = t.text; u->args = {parse_unary()};
            return u;
        }
        return parse_primary();
    }
    ExprP parse_if_expr() {
        DepthGuard guard(*this);
        const int pos = bump().pos;   // `if`
        ExprP c = parse_expr();
        ExprP a, b;
        if (take_punct("{")) { skip_nl_only(); a = parse_expr(); skip_nl_only(); expect_punct("}"); }
        else a = parse_expr();
        skip_nl_only();
        if (!at_ident("else")) {
            bool later_else = false;
            for (int k = 1; k < 64 && peek(k).kind != Tk::End; ++k) if (peek(k).kind == Tk::Ident && peek(k).text == "else") later_else = true;
            throw DslError(later_else ? "unexpected tokens after `if`/`else` expression" : "conditional expression needs an `else` branch", peek().pos);
        }
        bump();
        if (at_ident("if")) b = parse_if_expr();
        else if (take_punct("{")) { skip_nl_only(); b = parse_expr(); skip_nl_only(); expect_punct("}"); }
        else b = parse_expr();
        ExprP e = mk(Expr::IfElse, pos);
        e->args = {c, a, b};
        return e;
    }
    ExprP parse_primary() {
        const Token t = peek();
        if (t.kind == Tk::Number) { bump(); ExprP e = mk(Expr::Num, t.pos); e->num = t.num; return e; }
        if (t.kind == Tk::Punct && t.text == "(") {
            bump();
            ExprP e = parse_expr();
            expect_punct(")");
            return e;
        }
        if (t.kind == Tk::Ident) {
            if (t.text == "if") return parse_if_expr();
            bump();
            if (t.text == "true" || t.text == "false") { ExprP e = mk(Expr::BoolLit, t.pos); e->bval = t.text == "true"; return e; }
            if (at_punct("(")) {
                bump();
                ExprP c = mk(Expr::Call, t.pos);
                c->name = t.text;
                if (!at_punct(")")) {
                    do { c->args.push_back(parse_expr()); } while (take_punct(","));
                }
                expect_punct(")");
                return c;
            }
            if (at_punct("[")) {
                bump();
                ExprP ix = mk(Expr::Index, t.pos);
                ix->name = t.text;
                ix->args = {parse_expr()};
                expect_punct("]");
                return ix;
            }
            ExprP n = mk(Expr::Name, t.pos);
            n->name = t.text;
            return n;
        }
        throw DslError("expected an expression, found `" + t.text + "`", t.pos);
    }

    // ---- canonical statements (parser.rs:791-1020) ----------------------------------------------
    bool surface_if = false;     // authoring statement-level `if`: bodies hold plain-variable assignments only (authoring.rs:700-733)
    std::vector<Stmt> parse_stmt_body() {
        if (surface_if && !at_punct("{")) throw DslError("expected `{` to open `if`/`else` body", peek().pos);
        if (surface_if) {
            int depth = 0;
            bool closed = false;
            for (size_t k = 0; peek(k).kind != Tk::End; ++k) {
                if (peek(k).kind != Tk::Punct) continue;
                if (peek(k).text == "{") ++depth;
                else if (peek(k).text == "}" && --depth == 0) { closed = true; break; }
            }
            if (!closed) throw DslError("unclosed `{` in `if`/`else` body", peek().pos);
        }
        expect_punct("{");
        std::vector<Stmt> out;
        skip_newlines();
        while (!at_punct("}")) {
            if (peek().kind == Tk::End) throw DslError(surface_if ? "unclosed `{` in `if`/`else` body" : "unterminated block", peek().pos);
            if (surface_if && peek().kind == Tk::Ident && !at_ident("if") && peek(1).kind == Tk::Punct && peek(1).text == "(")
                throw DslError("an `if` statement body supports only plain-variable assignments (`x = <expression>`); to make `ddt(...)`, `out(...)`, or "
                               "`init(...)` conditional, use a conditional equation instead, e.g. `ddt(central) = if (cond) <a> else <b>`", peek().pos);
            if (surface_if && peek().kind == Tk::Ident && peek(1).kind == Tk::Punct && peek(1).text == "=" && (peek(2).kind == Tk::Newline || peek(2).kind == Tk::End || (peek(2).kind == Tk::Punct && peek(2).text == "}")))
                throw DslError("expected `name = <expression>`", peek().pos);
            out.push_back(parse_stmt());
            skip_newlines();
        }
        expect_punct("}");
        return out;
    }
    Stmt parse_stmt() {
        DepthGuard guard(*this);
        Stmt s;
        s.pos = peek().pos;
        if (at_ident("if")) {
            bump();
            s.kind = Stmt::If;
            s.cond = parse_expr();
            s.then_body = parse_stmt_body();
            skip_nl_only();
            if (at_ident("else")) {
                bump();
                if (at_ident("if")) s.else_body = {parse_stmt()};
                else s.else_body = parse_stmt_body();
            }
            return s;
        }
        if (at_ident("for")) {
            bump();
            s.kind = Stmt::For;
            s.var = expect_ident();
            if (!at_ident("in")) throw DslError("expected `in`", peek().pos);
            bump();
            s.lo = parse_expr();
            expect_punct("..");
            s.hi = parse_expr();
            s.body = parse_stmt_body();
            return s;
        }
        if (at_ident("let")) {
            bump();
            s.kind = Stmt::Let;
            s.target = expect_ident();
            expect_punct("=");
            s.value = parse_expr();
            return s;
        }
        s.kind = Stmt::Assign;
        // target: name | name[idx] | callee(place) ; output labels may be bare integers
        if (peek().kind == Tk::Number) {
            const Token t = bump();
            std::ostringstream os; os << (long long)t.num;
            s.target = os.str();
        } else {
            s.target = expect_ident();
        }
        if (take_punct("(")) {
            s.callee = s.target;
            if (peek().kind == Tk::Number) { std::ostringstream os; os << (long long)bump().num; s.target = os.str(); }
            else s.target = expect_ident();
            if (take_punct("[")) { s.index = parse_expr(); expect_punct("]"); }
            expect_punct(")");
        } else if (take_punct("[")) {
            s.index = parse_expr();
            expect_punct("]");
        }
        expect_punct("=");
        s.value = parse_expr();
        return s;
    }
};

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

ExprP parse_expr_str(const std::string& s, int base) {
    Parser p(lex(s, base, false));
    if (p.at_ident("if")) {      // a surface conditional spans the whole right-hand side (authoring.rs:1520-1543)
        ExprP e = p.parse_if_expr();
        if (p.peek().kind != Tk::End) throw DslError("unexpected tokens after `if`/`else` expression", p.peek().pos);
        return e;
    }
    ExprP e = p.parse_expr();
    if (p.peek().kind != Tk::End) throw DslError("unexpected trailing tokens `" + p.peek().text + "`", p.peek().pos);
    return e;
}

// find the first top-level occurrence (outside parentheses/brackets/braces) of `what`
size_t find_top_level(const std::string& s, const std::string& what, bool assignment) {
    int depth = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '(' || c == '[' || c == '{') ++depth;
        else if (c == ')' || c == ']' || c == '}') --depth;
        else if (depth == 0 && s.compare(i, what.size(), what) == 0) {
            if (assignment) {
                // a single '=' that is not part of ==, !=, <=, >=
                const char prev = i > 0 ? s[i - 1] : ' ';
                const char next = i + 1 < s.size() ? s[i + 1] : ' ';
                if (next == '=' || prev == '=' || prev == '!' || prev == '<' || prev == '>') continue;
            }
            return i;
        }
    }
    return std::string::npos;
}

std::vector<std::string> split_commas(const std::string& s) {
    std::vector<std::string> out;
    int depth = 0;
    std::string cur;
    for (char c : s) {
        if (c == '(' || c == '[') ++depth;
        if (c == ')' || c == ']') --depth;
        if (c == ',' && depth == 0) { out.push_back(trim(cur)); cur.clear(); }
        else cur.push_back(c);
    }
    if (!trim(cur).empty()) out.push_back(trim(cur));
    return out;
}

bool is_ident(const std::string& s) {
    if (s.empty() || !(std::isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
    for (char c : s) if (!(std::isalnum((unsigned char)c) || c == '_')) return false;
    return true;
}
bool is_label(const std::string& s) {   // identifier or non-negative integer (route / output labels)
    if (is_ident(s)) return true;
    if (s.empty()) return false;
    for (char c : s) if (!std::isdigit((unsigned char)c)) return false;
    return true;
}

void parse_place(const std::string& text, int base, std::string& name, ExprP& index) {
    const std::string t = trim(text);
    const size_t lb = t.find('[');
    if (lb == std::string::npos) {
        if (!is_ident(t)) throw DslError("expected a state name, found `" + t + "`", base);
        name = t; index = nullptr;
        return;
    }
    const size_t rb = t.rfind(']');
    if (rb == std::string::npos || rb < lb) throw DslError("expected `]`", base);
    name = trim(t.substr(0, lb));
    if (!is_ident(name)) throw DslError("expected a state name, found `" + name + "`", base);
    index = parse_expr_str(t.substr(lb + 1, rb - lb - 1), base + (int)lb + 1);
}

StateDecl parse_state_decl(const std::string& item, int base) {
    StateDecl d;
    const size_t lb = item.find('[');
    if (lb == std::string::npos) {
        if (!is_ident(item)) throw DslError("bad state declaration `" + item + "`", base);
        d.name = item;
        return d;
    }
    const size_t rb = item.find(']');
    d.name = trim(item.substr(0, lb));
    d.is_array = true;
    d.len = std::atoi(item.substr(lb + 1, rb - lb - 1).c_str());
    if (!is_ident(d.name) || d.len <= 0) throw DslError("bad state declaration `" + item + "`", base);
    return d;
}

// ---- authoring shorthand (authoring.rs:362-900) ------------------------------------------------------
ModelAst parse_authoring(const std::string& src) {
    ModelAst m;
    m.authoring = true;
    bool explicit_kind = false;
    std::map<std::string, std::pair<ExprP, ExprP>> route_mods;   // route -> (lag, fa)
    std::set<std::string> declared_outputs;
    std::vector<std::string> explicit_outputs, inferred_outputs;   // `outputs = ...` items; targets of out(...) seen before any declaration
    // ---- physical lines -> logical lines (authoring.rs: a statement continues while a `{` / `(` / `[` is open,
    // and an `else` that starts the next non-trivial line continues an `if`); comments are stripped per
    // physical line, offsets are kept for diagnostics
    struct Logical { std::string text; int base; };
    std::vector<Logical> logical;
    {
        std::vector<std::pair<std::string, int>> phys;
        size_t off0 = 0;
        while (off0 <= src.size()) {
            size_t eol = src.find('\n', off0);
            if (eol == std::string::npos) eol = src.size();
            std::string line = src.substr(off0, eol - off0);
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line = line.substr(0, hash);
            const size_t sl = line.find("//");
            if (sl != std::string::npos) line = line.substr(0, sl);
            phys.emplace_back(line, (int)off0);
            if (eol == src.size()) break;
            off0 = eol + 1;
        }
        auto starts_with_else = [](const std::string& l) {
            const std::string t = trim(l);
            return t.compare(0, 4, "else") == 0 && (t.size() == 4 || !(std::isalnum((unsigned char)t[4]) || t[4] == '_'));
        };
        size_t k = 0;
        while (k < phys.size()) {
            if (trim(phys[k].first).empty()) { ++k; continue; }
            Logical L{phys[k].first, phys[k].second};
            int depth = 0;
            auto scan = [&](const std::string& l) { for (char c : l) { if (c == '{' || c == '(' || c == '[') ++depth; else if (c == '}' || c == ')' || c == ']') --depth; } };
            scan(phys[k].first);
            ++k;
            while (k < phys.size()) {
                if (depth > 0) { L.text += "\n" + phys[k].first; scan(phys[k].first); ++k; continue; }
                size_t nx = k;
                while (nx < phys.size() && trim(phys[nx].first).empty()) ++nx;
                if (nx < phys.size() && starts_with_else(phys[nx].first)) {
                    for (; k <= nx; ++k) { L.text += "\n" + phys[k].first; scan(phys[k].first); }
                    continue;
                }
                break;
            }
            logical.push_back(L);
        }
    }
    for (size_t li = 0; li < logical.size(); ++li) {
        const std::string& line = logical[li].text;
        const int base = logical[li].base;
        const size_t eol = 0; (void)eol;
        const std::string t = trim(line);
        if (t.empty()) continue;

        if (t.compare(0, 2, "if") == 0 && (t.size() == 2 || !(std::isalnum((unsigned char)t[2]) || t[2] == '_'))) {
            Parser p(lex(t, base, true));
            p.surface_if = true;
            Stmt st = p.parse_stmt();
            p.skip_nl_only();
            if (p.peek().kind != Tk::End) throw DslError("unexpected tokens after `if`/`else` statement: `" + p.peek().text + "`", p.peek().pos);
            m.derive.push_back(st);
            continue;
        }
        const size_t arrow = find_top_level(t, "->", false);
        if (arrow != std::string::npos) {
            // bolus(route) -> place | infusion(route) -> place
            const std::string lhs = trim(t.substr(0, arrow)), rhs = trim(t.substr(arrow + 2));
            const size_t lp = lhs.find('('), rp = lhs.rfind(')');
            if (lp == std::string::npos || rp == std::string::npos) throw DslError("expected `bolus(route)` or `infusion(route)`", base);
            const std::string kind = trim(lhs.substr(0, lp)), label = trim(lhs.substr(lp + 1, rp - lp - 1));
            RouteDecl r;
            r.has_kind = true;
            if (kind == "bolus") r.kind = RouteKind::Bolus;
            else if (kind == "infusion") r.kind = RouteKind::Infusion;
            else throw DslError("unknown route shorthand `" + kind + "`", base);
            if (!is_label(label)) throw DslError("bad route label `" + label + "`", base);
            r.name = label;
            for (const auto& q : m.routes)
                if (q.name == r.name && q.kind == r.kind) throw DslError("duplicate route `" + r.name + "`", base);
            parse_place(rhs, base + (int)arrow + 2, r.dest, r.dest_index);
            m.routes.push_back(r);
            continue;
        }
        const size_t eq = find_top_level(t, "=", true);
        if (eq == std::string::npos) throw DslError("expected a declaration, equation, or route shorthand: `" + t + "`", base);
        const std::string lhs = trim(t.substr(0, eq));
        std::string rhs = trim(t.substr(eq + 1));
        const int rbase = base + (int)eq + 1;
        if (lhs == "name") { if (!is_ident(rhs)) throw DslError("expected `name = <identifier>`", base); m.name = rhs; }
        else if (lhs == "model") throw DslError("`model = ...` has been renamed to `name = ...`", base);
        else if (lhs == "kind") {
            if (rhs == "ode") m.kind = ModelKind::Ode;
            else if (rhs == "analytical") m.kind = ModelKind::Analytical;
            else if (rhs == "sde") m.kind = ModelKind::Sde;
            else throw DslError("unknown model kind `" + rhs + "`", base);
            explicit_kind = true;
        }
        else if (lhs == "params" || lhs == "parameters") { for (auto& s : split_commas(rhs)) { if (!is_ident(s)) throw DslError("bad parameter `" + s + "`", base); m.params.push_back(s); } }
        else if (lhs == "covariates") {
            for (auto& s : split_commas(rhs)) {
                CovDecl c;
                const size_t at = s.find('@');
                c.name = trim(at == std::string::npos ? s : s.substr(0, at));
                if (at != std::string::npos) c.interpolation = trim(s.substr(at + 1));
                if (!is_ident(c.name)) throw DslError("bad covariate `" + s + "`", base);
                m.covariates.push_back(c);
            }
        }
        else if (lhs == "states") { for (auto& s : split_commas(rhs)) m.states.push_back(parse_state_decl(s, base)); }
        else if (lhs == "derived") {
            for (auto& s : split_commas(rhs)) {
                for (auto& d : m.derived_decl) if (d == s) throw DslError("duplicate derived declaration `" + s + "`", base);
                m.derived_decl.push_back(s);
            }
        }
        else if (lhs == "outputs") {
            // an `outputs = ...` line that follows inferred `out(...)` targets still has to cover them
            for (auto& s : split_commas(rhs)) { if (!is_label(s)) throw DslError("bad output `" + s + "`", base); explicit_outputs.push_back(s); }
            for (auto& o : inferred_outputs) {
                bool ok = false;
                for (auto& e : explicit_outputs) ok = ok || e == o;
                if (!ok) throw DslError("output `" + o + "` is not declared in `outputs = ...`", base);
            }
            if (!inferred_outputs.empty()) { m.outputs_decl.clear(); declared_outputs.clear(); inferred_outputs.clear(); }
            for (auto& s : split_commas(rhs)) if (!declared_outputs.count(s)) { m.outputs_decl.push_back(s); declared_outputs.insert(s); }
        }
        else if (lhs == "particles") { m.particles = (int)std::strtol(rhs.c_str(), nullptr, 10); }
        else if (lhs == "function") throw DslError("`function = ...` has been renamed to `structure = ...`", base);
        else if (lhs == "structure") { m.structure = rhs; }
        else if (lhs.compare(0, 6, "const ") == 0) { m.constants.emplace_back(trim(lhs.substr(6)), parse_expr_str(rhs, rbase)); }
        else if (lhs.find('(') != std::string::npos) {
            const size_t lp = lhs.find('('), rp = lhs.rfind(')');
            if (rp == std::string::npos) throw DslError("expected `)` to close call-style target", base);
            const std::string callee = trim(lhs.substr(0, lp)), arg = trim(lhs.substr(lp + 1, rp - lp - 1));
            if (callee == "lag" || callee == "fa") {
                auto& mod = route_mods[arg];
                ExprP v = parse_expr_str(rhs, rbase);
                if (callee == "lag") { if (mod.first) throw DslError("duplicate route property `lag`", base); mod.first = v; }
                else { if (mod.second) throw DslError("duplicate route property `bioavailability`", base); mod.second = v; }
            } else if (callee == "dx" || callee == "ddt" || callee == "noise" || callee == "init") {
                Stmt s; s.kind = Stmt::Assign; s.pos = base;
                s.callee = (callee == "dx") ? "ddt" : callee;
                parse_place(arg, base + (int)lp + 1, s.target, s.index);
                s.value = parse_expr_str(rhs, rbase);
                if (callee == "noise") m.diffusion.push_back(s);
                else if (callee == "init") m.init.push_back(s);
                else m.dynamics.push_back(s);
            } else if (callee == "out") {
                if (!is_label(arg)) throw DslError("bad output label `" + arg + "`", base);
                if (arg.compare(0, 6, "input_") == 0 && arg.size() > 6 && std::all_of(arg.begin() + 6, arg.end(), [](char ch) { return std::isdigit((unsigned char)ch) != 0; }))
                    throw DslError("`" + arg + "` is a route label and cannot be used as an output; use `outeq_" + arg.substr(6) + "` here", base);
                if (!explicit_outputs.empty() && !declared_outputs.count(arg)) throw DslError("output `" + arg + "` is not declared in `outputs = ...`", base);
                if (!declared_outputs.count(arg)) { declared_outputs.insert(arg); m.outputs_decl.push_back(arg); inferred_outputs.push_back(arg); }
                const size_t tilde = find_top_level(rhs, "~", false);     // `~ continuous()` annotation
                if (tilde != std::string::npos) {
                    std::string ann = trim(rhs.substr(tilde + 1)), squeezed;
                    for (char ch : ann) if (!std::isspace((unsigned char)ch)) squeezed.push_back(ch);
                    if (squeezed != "continuous()") throw DslError("expected the output annotation `continuous()`, found `" + ann + "`", base);
                    rhs = trim(rhs.substr(0, tilde));
                }
                Stmt s; s.kind = Stmt::Assign; s.pos = base; s.callee = "out"; s.target = arg;
                s.value = parse_expr_str(rhs, rbase);
                m.outputs.push_back(s);
            } else {
                throw DslError("unsupported equation target `" + callee + "`", base);
            }
        }
        else {
            if (!is_label(lhs)) throw DslError("expected `name = <expression>`, found `" + lhs + "`", base);
            Stmt s; s.kind = Stmt::Assign; s.pos = base; s.target = lhs;
            s.value = parse_expr_str(rhs, rbase);
            if (declared_outputs.count(lhs)) { s.callee = "out"; m.outputs.push_back(s); }
            else m.derive.push_back(s);
        }
    }
    for (auto& kv : route_mods) {
        bool found = false;
        for (auto& r : m.routes) {
            if (r.name != kv.first) continue;
            if (r.kind == RouteKind::Infusion) {
                // lag / bioavailability are bolus-only; a same-named bolus route takes them
                bool has_bolus = false;
                for (auto& q : m.routes) if (q.name == kv.first && q.kind == RouteKind::Bolus) has_bolus = true;
                if (!has_bolus) throw DslError(std::string("DSL authoring does not allow `") + (kv.second.first ? "lag" : "bioavailability") + "` on infusion route `" + r.name + "`");
                continue;
            }
            r.lag = kv.second.first; r.fa = kv.second.second; found = true;
        }
        if (!found) throw DslError("route property refers to unknown route `" + kv.first + "`");
    }
    if (!explicit_kind) {
        // determine_kind (authoring.rs:898-): sde if noise/particles, analytical if structure, else ode
        if (!m.diffusion.empty() || m.particles > 0) m.kind = ModelKind::Sde;
        else if (!m.structure.empty()) m.kind = ModelKind::Analytical;
        else m.kind = ModelKind::Ode;
    }
    if (m.name.empty()) throw DslError("missing `name = <identifier>`");
    if (!m.derived_decl.empty()) {      // authoring.rs:1010-1027: with a `derived = ...` line every derive target must be listed
        std::vector<const Stmt*> stack;
        for (const auto& s : m.derive) stack.push_back(&s);
        while (!stack.empty()) {
            const Stmt* s = stack.back(); stack.pop_back();
            if (s->kind == Stmt::Assign && s->callee.empty() && std::find(m.derived_decl.begin(), m.derived_decl.end(), s->target) == m.derived_decl.end())
                throw DslError("derived value `" + s->target + "` is not declared in `derived = ...`", s->pos);
            for (const auto& q : s->then_body) stack.push_back(&q);
            for (const auto& q : s->else_body) stack.push_back(&q);
            for (const auto& q : s->body) stack.push_back(&q);
        }
    }
    return m;
}

// ---- canonical `model name { ... }` (parser.rs:300-790) ----------------------------------------------
ModelAst parse_canonical(const std::string& src) {
    Parser p(lex(src, 0, true));
    ModelAst m;
    p.skip_newlines();
    if (!p.at_ident("model")) throw DslError("expected `model`", p.peek().pos);
    p.bump();
    m.name = p.expect_ident();
    p.skip_nl_only();
    p.expect_punct("{");
    p.skip_newlines();
    bool have_kind = false;
    while (!p.at_punct("}")) {
        if (p.peek().kind == Tk::End) throw DslError("unexpected end of input in model body", p.peek().pos);
        const Token kw = p.bump();
        if (kw.kind != Tk::Ident) throw DslError("unexpected token `" + kw.text + "` in model body", kw.pos);
        const std::string k = kw.text;
        if (k == "kind") {
            const std::string v = p.expect_ident();
            if (v == "ode") m.kind = ModelKind::Ode;
            else if (v == "analytical") m.kind = ModelKind::Analytical;
            else if (v == "sde") m.kind = ModelKind::Sde;
            else throw DslError("expected `ode`, `analytical`, or `sde`", kw.pos);
            have_kind = true;
        } else if (k == "parameters") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            while (!p.at_punct("}")) { m.params.push_back(p.expect_ident()); p.skip_newlines(); }
            p.expect_punct("}");
        } else if (k == "constants") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            while (!p.at_punct("}")) {
                const std::string n = p.expect_ident();
                p.expect_punct("=");
                m.constants.emplace_back(n, p.parse_expr());
                p.skip_newlines();
            }
            p.expect_punct("}");
        } else if (k == "covariates") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            while (!p.at_punct("}")) {
                CovDecl c; c.name = p.expect_ident();
                if (p.take_punct("@")) c.interpolation = p.expect_ident();
                m.covariates.push_back(c);
                p.skip_newlines();
            }
            p.expect_punct("}");
        } else if (k == "states") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            while (!p.at_punct("}")) {
                StateDecl d; d.name = p.expect_ident();
                if (p.take_punct("[")) {
                    if (p.peek().kind != Tk::Number) throw DslError("state array size must be an integer constant", p.peek().pos);
                    d.len = (int)p.bump().num; d.is_array = true;
                    p.expect_punct("]");
                }
                m.states.push_back(d);
                p.skip_newlines();
            }
            p.expect_punct("}");
        } else if (k == "routes") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            while (!p.at_punct("}")) {
                RouteDecl r;
                // optional kind keyword, only when a label follows on the same line (parser.rs:611-640)
                if ((p.at_ident("bolus") || p.at_ident("infusion")) && (p.peek(1).kind == Tk::Ident || p.peek(1).kind == Tk::Number)) {
                    r.has_kind = true;
                    r.kind = p.bump().text == "bolus" ? RouteKind::Bolus : RouteKind::Infusion;
                }
                if (p.peek().kind == Tk::Number) { std::ostringstream os; os << (long long)p.bump().num; r.name = os.str(); }
                else r.name = p.expect_ident();
                p.expect_punct("->");
                r.dest = p.expect_ident();
                if (p.take_punct("[")) { r.dest_index = p.parse_expr(); p.expect_punct("]"); }
                if (p.at_punct("{")) {
                    p.bump(); p.skip_newlines();
                    while (!p.at_punct("}")) {
                        const std::string prop = p.expect_ident();
                        p.expect_punct("=");
                        ExprP v = p.parse_expr();
                        if (prop == "lag") r.lag = v;
                        else if (prop == "bioavailability") r.fa = v;
                        else throw DslError("unknown route property `" + prop + "`", kw.pos);
                        p.skip_newlines();
                    }
                    p.expect_punct("}");
                }
                m.routes.push_back(r);
                p.skip_newlines();
            }
            p.expect_punct("}");
        } else if (k == "derive" || k == "dynamics" || k == "outputs" || k == "init" || k == "drift" || k == "diffusion") {
            p.skip_nl_only();
            std::vector<Stmt> body = p.parse_stmt_body();
            if (k == "derive") m.derive = body;
            else if (k == "dynamics" || k == "drift") m.dynamics = body;
            else if (k == "outputs") { for (auto& s : body) if (s.kind == Stmt::Assign && s.callee.empty()) s.callee = "out"; m.outputs = body; }
            else if (k == "init") { for (auto& s : body) if (s.kind == Stmt::Assign && s.callee.empty()) s.callee = "init"; m.init = body; }
            else m.diffusion = body;
        } else if (k == "analytical") {
            p.skip_nl_only(); p.expect_punct("{"); p.skip_newlines();
            if (p.expect_ident() != "structure") throw DslError("expected `structure = <identifier>` inside analytical block", kw.pos);
            p.expect_punct("=");
            m.structure = p.expect_ident();
            p.skip_newlines();
            p.expect_punct("}");
        } else if (k == "particles") {
            if (p.peek().kind != Tk::Number) throw DslError("expected particle count", p.peek().pos);
            m.particles = (int)p.bump().num;
        } else {
            throw DslError("unexpected token `" + k + "` in model body", kw.pos);
        }
        p.skip_newlines();
    }
    p.expect_punct("}");
    if (!have_kind) throw DslError("model `" + m.name + "` does not declare a kind");
    // outputs of a canonical model = assignment targets of the outputs block, in order
    std::function<void(const std::vector<Stmt>&)> collect = [&](const std::vector<Stmt>& ss) {
        for (const auto& s : ss) {
            if (s.kind == Stmt::Assign) {
                bool seen = false;
                for (auto& o : m.outputs_decl) if (o == s.target) seen = true;
                if (!seen) m.outputs_decl.push_back(s.target);
            } else if (s.kind == Stmt::If) { collect(s.then_body); collect(s.else_body); }
            else if (s.kind == Stmt::For) collect(s.body);
        }
    };
    collect(m.outputs);
    return m;
}

}  // namespace

ModelAst parse_model(const std::string& source) {
    // canonical sources start with the `model` keyword followed by an identifier and `{`
    size_t i = 0;
    while (i < source.size()) {
        if (std::isspace((unsigned char)source[i])) { ++i; continue; }
        if (source[i] == '#') { while (i < source.size() && source[i] != '\n') ++i; continue; }
        break;
    }
    if (source.compare(i, 5, "model") == 0 && i + 5 < source.size() && std::isspace((unsigned char)source[i + 5])) {
        size_t j = i + 5;
        while (j < source.size() && std::isspace((unsigned char)source[j])) ++j;
        if (j < source.size() && (std::isalpha((unsigned char)source[j]) || source[j] == '_')) return parse_canonical(source);
    }
    return parse_authoring(source);
}

}  // namespace dsl
}  // namespace pharmsol
