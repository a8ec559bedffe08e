// Minimal JSON reader for solver parameters (objects, arrays, strings, numbers, bools, null).
// The polysolve side serialises its nlohmann::json with dump() and passes the text across the C ABI.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace psb {

struct JValue
{
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JValue> arr;
    std::map<std::string, JValue> obj;

    bool is_obj() const { return kind == Obj; }
    bool contains(const std::string &k) const { return kind == Obj && obj.count(k); }
    const JValue &at(const std::string &k) const
    {
        auto it = obj.find(k);
        if (it == obj.end())
            throw std::runtime_error("json: missing key " + k);
        return it->second;
    }
    double as_num() const
    {
        if (kind == Num) return num;
        if (kind == Bool) return b ? 1 : 0;
        throw std::runtime_error("json: number expected");
    }
    bool as_bool() const
    {
        if (kind == Bool) return b;
        if (kind == Num) return num != 0;
        throw std::runtime_error("json: bool expected");
    }
    const std::string &as_str() const
    {
        if (kind != Str) throw std::runtime_error("json: string expected");
        return str;
    }
};

class JParser
{
    const char *p, *e;
    int depth = 0; // nesting of the value being parsed: bounded, a malformed document must not overflow the stack
    static constexpr int kMaxDepth = 128;
    void ws() { while (p < e && std::isspace((unsigned char)*p)) ++p; }
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("json parse error: ") + m); }
    std::string str()
    {
        if (*p != '"') fail("expected string");
        ++p;
        std::string s;
        while (p < e && *p != '"')
        {
            if (*p == '\\' && p + 1 < e)
            {
                ++p;
                switch (*p)
                {
                case 'n': s += '\n'; break;
                case 't': s += '\t'; break;
                case 'r': s += '\r'; break;
                case 'b': s += '\b'; break;
                case 'f': s += '\f'; break;
                case 'u':
                    if (p + 4 < e) { s += '?'; p += 4; }
                    break;
                default: s += *p;
                }
                ++p;
            }
            else
                s += *p++;
        }
        if (p >= e) fail("unterminated string");
        ++p;
        return s;
    }
    struct Nest
    {
        JParser &P;
        explicit Nest(JParser &P_) : P(P_)
        {
            if (++P.depth > kMaxDepth) P.fail("nesting too deep");
        }
        ~Nest() { --P.depth; }
    };
    JValue val()
    {
        ws();
        if (p >= e) fail("unexpected end");
        JValue v;
        if (*p == '{')
        {
            Nest nest(*this);
            v.kind = JValue::Obj;
            ++p; ws();
            if (*p == '}') { ++p; return v; }
            for (;;)
            {
                ws();
                std::string k = str();
                ws();
                if (*p != ':') fail("expected ':'");
                ++p;
                v.obj[k] = val();
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
        }
        else if (*p == '[')
        {
            Nest nest(*this);
            v.kind = JValue::Arr;
            ++p; ws();
            if (*p == ']') { ++p; return v; }
            for (;;)
            {
                v.arr.push_back(val());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
        }
        else if (*p == '"')
        {
            v.kind = JValue::Str;
            v.str = str();
        }
        else if (!std::strncmp(p, "true", 4)) { v.kind = JValue::Bool; v.b = true; p += 4; }
        else if (!std::strncmp(p, "false", 5)) { v.kind = JValue::Bool; v.b = false; p += 5; }
        else if (!std::strncmp(p, "null", 4)) { v.kind = JValue::Null; p += 4; }
        else
        {
            char *end = nullptr;
            v.num = std::strtod(p, &end);
            if (end == p) fail("unexpected token");
            v.kind = JValue::Num;
            p = end;
        }
        return v;
    }

public:
    static JValue parse(const std::string &s)
    {
        JParser P;
        P.p = s.data();
        P.e = s.data() + s.size();
        JValue v = P.val();
        P.ws();
        if (P.p != P.e) P.fail("trailing characters");
        return v;
    }
};

inline std::string jnum(double v)
{
    if (!std::isfinite(v)) return "null";
    char buf[40];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    return buf;
}
inline std::string jstr(const std::string &s)
{
    std::string o = "\"";
    for (char c : s)
    {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else if ((unsigned char)c < 0x20)
        {
            char buf[8];
            std::snprintf(buf, sizeof buf, "\\u%04x", (unsigned)(unsigned char)c);
            o += buf;
        }
        else o += c;
    }
    return o + "\"";
}

} // namespace psb
