// json_min.h — minimal JSON reader for the saved-graph format (runtime.rs:44-48).  Numbers are f64,
// objects keep insertion order.  No external dependencies.
#pragma once
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

namespace jsonmin {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;
    bool is_number() const { return kind == Number; }
    bool is_string() const { return kind == String; }
    bool is_array() const { return kind == Array; }
    bool is_object() const { return kind == Object; }
    const Value* get(const char* key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const char* p;
    std::string err;
    void ws() { while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool str(std::string& out) {
        if (*p != '"') return fail("expected string");
        p++;
        while (*p && *p != '"') {
            if (*p == '\\') {
                p++;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {  // keep BMP code points as UTF-8
                        unsigned cp = 0;
                        for (int i = 0; i < 4; i++) {
                            p++;
                            char c = *p;
                            cp <<= 4;
                            if (c >= '0' && c <= '9') cp |= c - '0';
                            else if (c >= 'a' && c <= 'f') cp |= c - 'a' + 10;
                            else if (c >= 'A' && c <= 'F') cp |= c - 'A' + 10;
                            else return fail("bad \\u escape");
                        }
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    } break;
                    case 0: return fail("unterminated string");
                    default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (*p != '"') return fail("unterminated string");
        p++;
        return true;
    }
    bool value(Value& v, int depth) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (*p == '{') {
            v.kind = Value::Object;
            p++;
            ws();
            if (*p == '}') { p++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (*p != ':') return fail("expected ':'");
                p++;
                Value c;
                if (!value(c, depth + 1)) return false;
                v.obj.emplace_back(std::move(k), std::move(c));
                ws();
                if (*p == ',') { p++; continue; }
                if (*p == '}') { p++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = Value::Array;
            p++;
            ws();
            if (*p == ']') { p++; return true; }
            for (;;) {
                Value c;
                if (!value(c, depth + 1)) return false;
                v.arr.push_back(std::move(c));
                ws();
                if (*p == ',') { p++; continue; }
                if (*p == ']') { p++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = Value::String; return str(v.str); }
        if (!__builtin_strncmp(p, "true", 4)) { v.kind = Value::Bool; v.b = true; p += 4; return true; }
        if (!__builtin_strncmp(p, "false", 5)) { v.kind = Value::Bool; v.b = false; p += 5; return true; }
        if (!__builtin_strncmp(p, "null", 4)) { v.kind = Value::Null; p += 4; return true; }
        char* end = nullptr;
        double d = std::strtod(p, &end);
        if (end == p) return fail("unexpected character");
        v.kind = Value::Number;
        v.num = d;
        p = end;
        return true;
    }
};

inline bool parse(const char* text, Value& out, std::string& err) {
    Parser ps{text, {}};
    if (!ps.value(out, 0)) { err = ps.err; return false; }
    ps.ws();
    if (*ps.p) { err = "trailing characters"; return false; }
    return true;
}

}  // namespace jsonmin
