// literal.cpp — host side of the date/time slice: the text of a literal becomes the DATETIME / TIMESTAMP / DATE / TIME image the
// device compares against.  The reference parses with sscanf (str_to_datetime_internal, src/common/datetime.cpp:149-263;
// str_to_time, :477-560); this file walks the text with a small field scanner that accepts and rejects exactly what those format
// strings accept and reject (width-limited unsigned fields, "one or more characters outside [0-9a-z]" separators, stop at the
// first field that does not match), so the images agree on every input, malformed ones included.  tests/test_datetime.py holds
// the reference's own vectors (test/test_date_time.cpp) and a fuzz against the oracle's sscanf restatement.
#include <string.h>
#include <string>
#include "../include/bkgpu.h"
#include "datetime.h"
#include "plan.h"

namespace bk {
namespace {

inline bool digit(char c) { return c >= '0' && c <= '9'; }
inline bool c_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// A cursor over NUL-terminated text with sscanf's conversion rules.
struct Scan {
    const char* p;
    bool ok = true;
    // %<width>lu / %<width>u: optional blanks, optional sign (counts against the width), then at least one digit
    uint64_t uint_field(int width) {
        if (!ok) return 0;
        while (c_space(*p)) p++;
        const char* q = p; int w = width; bool neg = false;
        if (w > 0 && (*q == '+' || *q == '-')) { neg = *q == '-'; q++; w--; }
        if (w <= 0 || !digit(*q)) { ok = false; return 0; }
        uint64_t v = 0;
        while (w > 0 && digit(*q)) { v = v * 10 + (uint64_t)(*q - '0'); q++; w--; }
        p = q;
        return neg ? (uint64_t)0 - v : v;
    }
    // %d / strtoll-then-int: unbounded width; the long value saturates, then truncates to int the way the C conversion does
    int64_t int_field() {
        if (!ok) return 0;
        while (c_space(*p)) p++;
        const char* q = p; bool neg = false;
        if (*q == '+' || *q == '-') { neg = *q == '-'; q++; }
        if (!digit(*q)) { ok = false; return 0; }
        uint64_t v = 0; bool sat = false;
        const uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
        while (digit(*q)) {
            const uint64_t d = (uint64_t)(*q - '0');
            if (sat || v > (lim - d) / 10) sat = true; else v = v * 10 + d;
            q++;
        }
        p = q;
        if (sat) v = lim;
        return (int64_t)(int32_t)(uint32_t)(neg ? (uint64_t)0 - v : v);
    }
    // %*[^0-9a-z]: one or more characters that are neither digits nor lower-case letters
    void separator() {
        if (!ok) return;
        const char* q = p;
        while (*q && !digit(*q) && !(*q >= 'a' && *q <= 'z')) q++;
        if (q == p) ok = false;
        p = q;
    }
    // an ordinary format character: blanks in the format skip any blanks, anything else must match exactly
    void literal(char c) {
        if (!ok) return;
        if (c == ' ') { while (c_space(*p)) p++; return; }
        if (*p != c) { ok = false; return; }
        p++;
    }
};

}  // namespace

// str_to_datetime_internal (datetime.cpp:149-263).  *is_full: the text carried a time of day (drives str_to_time's choice).
uint64_t parse_datetime(const char* text, size_t length, bool* is_full) {
    size_t lead = 0;
    while (lead < length && text[lead] == ' ') lead++;
    text += lead;
    constexpr size_t CAP = 26;
    const size_t len = length < CAP ? length : CAP;      // the reference keeps the leading blanks in this bound (datetime.cpp:155-159)
    char buf[CAP + 1] = {0};
    size_t real = 0;
    while (real < len && real < length - lead && text[real]) real++;   // a C string: an embedded NUL ends it
    memcpy(buf, text, real);
    bool delimited = !(digit(buf[2]) && digit(buf[4]));
    if (buf[3] == '-') delimited = true;                // YYY-MM-DD
    int year_digits = -1, seps = 0;
    size_t dot = 0;
    for (; dot < len; dot++) {
        if (delimited) {
            if (!digit(buf[dot])) { seps++; if (year_digits < 0) year_digits = (int)dot; }
            if (seps > 5 && buf[dot] == '.') break;
        } else if (buf[dot] == '.') break;
    }
    if (dot < len) for (size_t i = dot + 1; i <= dot + 6 && i < CAP; i++) if (!digit(buf[i])) buf[i] = '0';   // ".5" reads as 500000 microseconds
    uint64_t f[7] = {0, 0, 0, 0, 0, 0, 0};   // year month day hour minute second microsecond
    bool full = false;
    Scan s{buf};
    auto run = [&](int n_fields, const int* widths, bool with_seps, bool micro) {
        for (int i = 0; i < n_fields && s.ok; i++) {
            if (i > 0 && with_seps) s.separator();
            const uint64_t v = s.uint_field(widths[i]);
            if (s.ok) f[i] = v;
        }
        if (micro && s.ok) { s.literal('.'); const uint64_t v = s.uint_field(6); if (s.ok) f[6] = v; }
    };
    static const int W4[6] = {4, 2, 2, 2, 2, 2}, W2[6] = {2, 2, 2, 2, 2, 2};
    if (delimited) { run(6, W4, true, true); full = true; }
    else if (dot <= 6) { run(3, W2, false, false); year_digits = 2; }
    else if (dot == 8) run(3, W4, false, false);
    else if (dot == 12) { run(6, W2, false, true); full = true; year_digits = 2; }
    else if (dot <= 13) { run(6, W2, false, false); full = true; year_digits = 2; }
    else if (dot >= 14) { run(6, W4, false, true); full = true; }
    else return 0;   // 7, 9, 10, 11 digits: no layout
    if (year_digits == 2) { if (f[0] >= 70 && f[0] < 100) f[0] += 1900; else if (f[0] < 70 && f[0] > 0) f[0] += 2000; }
    if (f[1] > 12 || f[2] > 31 || f[3] > 23 || f[4] > 59 || f[5] > 59) return 0;
    if (is_full) *is_full = full;
    return dt_make(f[0], f[1], f[2], f[3], f[4], f[5], f[6]);
}

// str_to_time (datetime.cpp:477-560): "[-][D ]H:M:S", "[-]HHMMSS" (right-aligned), or a full date-time whose time of day is kept
int32_t parse_time(const char* text, size_t length) {
    while (length > 0 && *text == ' ') { text++; length--; }
    bool minus = false;
    if (length > 0 && *text == '-') { minus = true; text++; length--; }
    const size_t len = length < 20 ? length : 20;
    bool blank = false, delim = false;
    size_t dot = 0;
    for (; dot < len; dot++) {
        const char c = text[dot];
        if (c == ' ') blank = delim = true;
        if (c == ':') delim = true;
        if (c == '.') break;
    }
    if (dot >= 12) {   // long enough for YYMMDDHHMMSS: a date-time, if it parses as one
        bool full = false;
        const uint64_t dt = parse_datetime(text, length, &full);
        if (full) return (int32_t)dt_datetime_to_time(dt);
    }
    int64_t day = 0, hour = 0, minute = 0, second = 0;
    std::string z(text, strnlen(text, length));   // the scanner wants a terminator
    Scan s{z.c_str()};
    if (blank) {            // "%d %u:%2u:%2u"
        day = s.int_field(); s.literal(' ');
        { const uint64_t v = s.uint_field(1 << 20); if (s.ok) hour = (int64_t)(int32_t)(uint32_t)v; }
        s.literal(':'); { const uint64_t v = s.uint_field(2); if (s.ok) minute = (int64_t)v; }
        s.literal(':'); { const uint64_t v = s.uint_field(2); if (s.ok) second = (int64_t)v; }
    } else if (delim) {     // "%d:%2u:%2u"
        hour = s.int_field();
        s.literal(':'); { const uint64_t v = s.uint_field(2); if (s.ok) minute = (int64_t)v; }
        s.literal(':'); { const uint64_t v = s.uint_field(2); if (s.ok) second = (int64_t)v; }
    } else {                // digits only: seconds, then minutes, then hours, from the right (strtoll on each piece)
        auto piece = [&](size_t from, size_t n) -> int64_t {
            Scan t{nullptr}; std::string sub(text + from, n); t.p = sub.c_str();
            const int64_t v = t.int_field(); return t.ok ? v : 0;
        };
        size_t i = dot;
        if (i >= 4) { second = piece(i - 2, 2); minute = piece(i - 4, 2); hour = piece(0, i - 4); }
        else if (i >= 2) { second = piece(i - 2, 2); minute = piece(0, i - 2); }
        else second = piece(0, i);
    }
    if (day < 0 || hour < 0 || minute < 0 || minute > 59 || second < 0 || second > 59) return 0;
    hour += day * 24;
    const int32_t t = (int32_t)((uint32_t)second | ((uint32_t)minute << 6) | ((uint32_t)hour << 12));
    return minus ? -t : t;
}

// ExprValue::cast_to with a STRING source and a date/time target (expr_value.h:534-573)
uint64_t parse_literal(const char* text, size_t length, int to_prim) {
    switch (to_prim) {
        case BK_DATETIME: return parse_datetime(text, length, nullptr);
        case BK_TIMESTAMP: return (uint64_t)(uint32_t)dt_datetime_to_timestamp(parse_datetime(text, length, nullptr));
        case BK_DATE: return dt_datetime_to_date(parse_datetime(text, length, nullptr));
        default: return (uint64_t)(int64_t)parse_time(text, length);
    }
}

}  // namespace bk

// The same parser for callers that hold the text of a literal (a reference-side binding folding `col >= '2024-01-01'` before it
// builds the plan words): `prim_type` is BK_DATETIME / BK_TIMESTAMP / BK_DATE / BK_TIME, *image the canonical 64-bit image.
extern "C" int bkgpu_parse_datetime(const char* text, size_t length, int prim_type, uint64_t* image) {
    if (!text || !image || !bk::dt_is_family(prim_type)) return BKGPU_EINVAL;
    *image = bk::parse_literal(text, length, prim_type);
    return BKGPU_OK;
}

// ExprValue::cast_to between two non-STRING types on canonical images (what plan compilation applies to literals): for a binding
// that folds constants, and for tests of the calendar arithmetic.  A TIME source with another date/time target is refused (it is
// relative to the current date in the reference).
extern "C" int bkgpu_cast_image(uint64_t image, int from_prim, int to_prim, uint64_t* out) {
    auto known = [](int t) { return t >= BK_BOOL && t <= BK_TIME && t != BK_STRING && t != BK_HLL; };
    if (!out || !known(from_prim) || !known(to_prim)) return BKGPU_EINVAL;
    if (from_prim == BK_TIME && to_prim != BK_TIME && bk::dt_is_family(to_prim)) return BKGPU_EUNSUPPORTED;
    *out = bk::host_cast_prim(image, from_prim, to_prim);
    return BKGPU_OK;
}
