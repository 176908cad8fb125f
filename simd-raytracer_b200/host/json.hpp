// host/json.hpp - a small strict JSON DOM reader, enough for .crtscene files.
//
// The reference parses with simdjson v3.13.0 (CMakeLists.txt:10-15), which is not available offline.  Numbers are
// converted with std::from_chars, i.e. correctly rounded and locale independent, so every double equals simdjson's
// and float(double) reproduces io/json/loader.hpp:9-17 exactly.
#pragma once

#include <charconv>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace rtb::json {

struct parse_error : std::runtime_error { using std::runtime_error::runtime_error; };

struct value;
using array = std::vector<value>;
using object = std::vector<std::pair<std::string, value>>;   // insertion order kept; first match wins on lookup

struct value {
    enum kind_t { NUL, BOOL, NUMBER, STRING, ARRAY, OBJECT } kind = NUL;
    bool b = false;
    double num = 0;
    bool is_integer = false;         // written without fraction / exponent and non-negative: a simdjson uint64
    uint64_t u64 = 0;
    std::string str;
    std::shared_ptr<array> arr;
    std::shared_ptr<object> obj;

    bool is_array() const { return kind == ARRAY; }
    bool is_object() const { return kind == OBJECT; }
    bool is_string() const { return kind == STRING; }
    bool is_number() const { return kind == NUMBER; }

    const value* find(std::string_view key) const {
        if (kind != OBJECT) return nullptr;
        for (const auto& kv : *obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const value& at(std::string_view key) const {
        const value* v = find(key);
        if (!v) throw parse_error("missing key \"" + std::string(key) + "\"");
        return *v;
    }
    const array& as_array() const { if (kind != ARRAY) throw parse_error("expected array"); return *arr; }
    double as_double() const { if (kind != NUMBER) throw parse_error("expected number"); return num; }
    uint64_t as_u64() const { if (kind != NUMBER || !is_integer) throw parse_error("expected unsigned integer"); return u64; }
    bool as_bool() const { if (kind != BOOL) throw parse_error("expected bool"); return b; }
    const std::string& as_string() const { if (kind != STRING) throw parse_error("expected string"); return str; }
};

class parser {
public:
    explicit parser(std::string_view text) : p_(text.data()), end_(text.data() + text.size()) {}

    value parse_document() {
        skip_ws();
        value v = parse_value(0);
        skip_ws();
        if (p_ != end_) fail("trailing characters after JSON document");
        return v;
    }

private:
    const char* p_;
    const char* end_;

    [[noreturn]] void fail(const char* msg) const { throw parse_error(std::string("JSON: ") + msg); }
    void skip_ws() { while (p_ != end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_; }
    bool consume(char c) { if (p_ != end_ && *p_ == c) { ++p_; return true; } return false; }
    void expect_word(const char* w) {
        const std::size_t n = std::strlen(w);
        if (std::size_t(end_ - p_) < n || std::memcmp(p_, w, n) != 0) fail("bad literal");
        p_ += n;
    }

    value parse_value(int depth) {
        if (depth > 64) fail("nesting too deep");
        if (p_ == end_) fail("unexpected end");
        value v;
        switch (*p_) {
            case '{': {
                ++p_;
                v.kind = value::OBJECT;
                v.obj = std::make_shared<object>();
                skip_ws();
                if (consume('}')) return v;
                for (;;) {
                    skip_ws();
                    if (p_ == end_ || *p_ != '"') fail("expected object key");
                    std::string k = parse_string();
                    skip_ws();
                    if (!consume(':')) fail("expected ':'");
                    skip_ws();
                    v.obj->emplace_back(std::move(k), parse_value(depth + 1));
                    skip_ws();
                    if (consume(',')) continue;
                    if (consume('}')) return v;
                    fail("expected ',' or '}'");
                }
            }
            case '[': {
                ++p_;
                v.kind = value::ARRAY;
                v.arr = std::make_shared<array>();
                skip_ws();
                if (consume(']')) return v;
                for (;;) {
                    skip_ws();
                    v.arr->push_back(parse_value(depth + 1));
                    skip_ws();
                    if (consume(',')) continue;
                    if (consume(']')) return v;
                    fail("expected ',' or ']'");
                }
            }
            case '"': v.kind = value::STRING; v.str = parse_string(); return v;
            case 't': expect_word("true"); v.kind = value::BOOL; v.b = true; return v;
            case 'f': expect_word("false"); v.kind = value::BOOL; v.b = false; return v;
            case 'n': expect_word("null"); v.kind = value::NUL; return v;
            default: return parse_number();
        }
    }

    value parse_number() {
        const char* s = p_;
        bool integral = true;
        if (p_ != end_ && *p_ == '-') { ++p_; integral = false; }
        if (p_ == end_ || *p_ < '0' || *p_ > '9') fail("bad number");
        while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        if (p_ != end_ && *p_ == '.') {
            integral = false; ++p_;
            if (p_ == end_ || *p_ < '0' || *p_ > '9') fail("bad fraction");
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ != end_ && (*p_ == 'e' || *p_ == 'E')) {
            integral = false; ++p_;
            if (p_ != end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ == end_ || *p_ < '0' || *p_ > '9') fail("bad exponent");
            while (p_ != end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        value v;
        v.kind = value::NUMBER;
        auto r = std::from_chars(s, p_, v.num);
        if (r.ec != std::errc() || r.ptr != p_) fail("number out of range");
        if (integral) {
            auto ri = std::from_chars(s, p_, v.u64);
            v.is_integer = ri.ec == std::errc() && ri.ptr == p_;
        }
        return v;
    }

    static void append_utf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) out.push_back(char(cp));
        else if (cp < 0x800) { out.push_back(char(0xC0 | (cp >> 6))); out.push_back(char(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back(char(0xE0 | (cp >> 12))); out.push_back(char(0x80 | ((cp >> 6) & 0x3F))); out.push_back(char(0x80 | (cp & 0x3F))); }
        else { out.push_back(char(0xF0 | (cp >> 18))); out.push_back(char(0x80 | ((cp >> 12) & 0x3F))); out.push_back(char(0x80 | ((cp >> 6) & 0x3F))); out.push_back(char(0x80 | (cp & 0x3F))); }
    }
    uint32_t parse_hex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        ++p_;  // opening quote
        std::string out;
        for (;;) {
            if (p_ == end_) fail("unterminated string");
            const char c = *p_++;
            if (c == '"') return out;
            if (static_cast<unsigned char>(c) < 0x20) fail("control character in string");
            if (c != '\\') { out.push_back(c); continue; }
            if (p_ == end_) fail("unterminated escape");
            const char e = *p_++;
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = parse_hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        const uint32_t lo = parse_hex4();
                        if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else fail("bad surrogate pair");
                    }
                    append_utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
};

inline value parse(std::string_view text) { return parser(text).parse_document(); }

}  // namespace rtb::json
