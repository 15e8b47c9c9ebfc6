/*
 * bk_oracle.c — CPU restatement of the reference ROW ENGINE for the hot path
 * (scan -> filter -> aggregate / hash join / sort).  TEST INFRASTRUCTURE ONLY: see the
 * header of bk_oracle.h for who may load it and for the parity-pin status.
 *
 * It follows, function by function (all paths under /root/reference):
 *   ExprValue get_numberic/cast_to/add/compare/compare_diff_type
 *                                    include/common/expr_value.h:340-410,502-611,840-989
 *   MutTableKey / KeyEncoder         include/common/mut_table_key.h:60-200, key_encoder.h:120-170
 *   SlotRef / Literal get_value      include/expr/slot_ref.h:31-40, include/expr/literal.h:204-206
 *   ScalarFnCall type_inferer/get_value   src/expr/scalar_fn_call.cpp:40-120,194-225
 *   FunctionManager::complete_fn     src/expr/fn_manager.cpp:316-409
 *   operators                        src/expr/operators.cpp:18-103
 *   And/Or/Not/Xor/IsNull/IsTrue/In  include/expr/predicate.h:25-345, src/expr/predicate.cpp:102-189
 *   FilterNode::need_copy/get_next   src/exec/filter_node.cpp:726-795
 *   ExecNode::encode_exprs_key       src/exec/exec_node.cpp:555-571
 *   AggNode open/process_row_batch/get_next   src/exec/agg_node.cpp:405-573
 *   AggFnCall initialize/update/merge/finalize src/expr/agg_fn_call.cpp:370-410,496-555,719-822,927-990
 *   MemRow get/set_value storage casts        include/common/message_helper.h:155-253
 *   SortNode / Sorter / TopNSorter / MemRowCompare  src/exec/sort_node.cpp:278-385,
 *        src/runtime/sorter.cpp:54-114, src/runtime/topn_sorter.cpp:25-103,
 *        include/runtime/topn_sorter.h:96-106, src/mem_row/mem_row_compare.cpp:18-38
 *   Joiner encode_hash_key/construct_hash_map/construct_result_batch src/exec/joiner.cpp:166-217,608-685
 *   JoinNode::get_next_for_hash_inner_join    src/exec/join_node.cpp:1277-1326
 *   LimitNode                                 src/exec/limit_node.cpp:21-134
 *
 * Deliberately the same shape as the reference: pull-based get_next over row batches
 * of <= 1024 rows, one ExprValue per value access, byte-string group keys.
 */
#include "bk_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- enum values: the reference's (see include/bkgpu_plan.h for the citations) ---- */
enum { T_INVALID = 0, T_NULL = 1, T_BOOL = 2, T_INT8 = 3, T_INT16 = 4, T_INT32 = 5, T_INT64 = 6,
       T_UINT8 = 7, T_UINT16 = 8, T_UINT32 = 9, T_UINT64 = 10, T_FLOAT = 11, T_DOUBLE = 12,
       T_STRING = 13, T_DATETIME = 14, T_TIMESTAMP = 15, T_DATE = 16, T_HLL = 17, T_TIME = 18,
       T_MAXVALUE = 24 };
enum { N_SCAN = 1, N_SORT = 2, N_AGG = 4, N_MERGE_AGG = 5, N_TABLE_FILTER = 6, N_JOIN = 7,
       N_LIMIT = 11, N_WHERE_FILTER = 12, N_HAVING_FILTER = 13, N_PACKET = 14, N_SELECT_MANAGER = 25 };
enum { E_SLOT_REF = 1, E_FUNCTION_CALL = 2, E_AGG_EXPR = 3, E_NULL_LITERAL = 4, E_BOOL_LITERAL = 5,
       E_INT_LITERAL = 6, E_DOUBLE_LITERAL = 7, E_STRING_LITERAL = 8, E_IS_NULL = 9, E_IN = 10,
       E_LIKE = 11, E_NOT = 12, E_AND = 13, E_OR = 14, E_XOR = 15, E_TIMESTAMP_LITERAL = 16, E_DATETIME_LITERAL = 17,
       E_DATE_LITERAL = 18, E_IS_TRUE = 19, E_TIME_LITERAL = 20, E_ROW_EXPR = 22 };
enum { FT_COMMON = 0, FT_AGG = 1, FT_BIT_NOT = 2, FT_LOGIC_NOT = 3, FT_UMINUS = 4, FT_ADD = 5,
       FT_MINUS = 6, FT_MULTIPLIES = 7, FT_DIVIDES = 8, FT_MOD = 9, FT_LS = 10, FT_RS = 11,
       FT_BIT_AND = 12, FT_BIT_OR = 13, FT_BIT_XOR = 14, FT_EQ = 15, FT_NE = 16, FT_GT = 17,
       FT_GE = 18, FT_LT = 19, FT_LE = 20, FT_LOGIC_AND = 21, FT_LOGIC_OR = 22, FT_LOGIC_XOR = 23,
       FT_IS_NULL = 24, FT_IS_TRUE = 25, FT_IS_UNKNOWN = 26, FT_IN = 27, FT_LIKE = 28 };
enum { J_LEFT = 1, J_RIGHT = 2, J_INNER = 3, J_SEMI = 4, J_ANTI = 5 };
enum { A_COUNT_STAR, A_COUNT, A_SUM, A_AVG, A_MIN, A_MAX };

#define MAX_TUPLES 8
#define ROW_BATCH_CAPACITY 1024 /* include/common/common.h:162 */

/* =========================== ExprValue =========================== */
typedef struct { double sum; int64_t count; } AvgIntermediate; /* include/expr/agg_fn_call.h:41-47 */
typedef struct ExprValue {
    int type;
    union {
        uint8_t bool_val; int8_t int8_val; int16_t int16_val; int32_t int32_val; int64_t int64_val;
        uint8_t uint8_val; uint16_t uint16_val; uint32_t uint32_val; uint64_t uint64_val;
        float float_val; double double_val;
    } u;
    AvgIntermediate avg; /* the 16-byte STRING blob AVG keeps in its intermediate slot */
    const char* str;     /* STRING literal (NUL-terminated, owned by its Expr); NULL for the AVG blob */
} ExprValue;

static ExprValue ev_null(void) { ExprValue v; memset(&v, 0, sizeof v); v.type = T_NULL; return v; }
static ExprValue ev_typed(int t) { ExprValue v; memset(&v, 0, sizeof v); v.type = t; return v; }
static ExprValue ev_bool(int b) { ExprValue v = ev_typed(T_BOOL); v.u.bool_val = b ? 1 : 0; return v; }
static int ev_is_null(const ExprValue* v) { return v->type == T_NULL || v->type == T_INVALID; }

static int is_int(int t) { return t >= T_INT8 && t <= T_UINT64; }
static int is_uint(int t) { return t >= T_UINT8 && t <= T_UINT64; }
static int is_signed(int t) { return t >= T_INT8 && t <= T_INT64; }
static int is_double(int t) { return t == T_FLOAT || t == T_DOUBLE; }
static int is_string(int t) { return t == T_STRING || t == T_HLL; }

/* get_numberic<T>: C++ static_cast chains, expr_value.h:340-410 (float_precision_len = -1) */
#define GET_NUM(NAME, CT)                                                              \
    static CT NAME(const ExprValue* v) {                                               \
        switch (v->type) {                                                             \
            case T_BOOL: return (CT)v->u.bool_val;                                     \
            case T_INT8: return (CT)v->u.int8_val;                                     \
            case T_INT16: return (CT)v->u.int16_val;                                   \
            case T_INT32: case T_TIME: return (CT)v->u.int32_val;                      \
            case T_INT64: return (CT)v->u.int64_val;                                   \
            case T_UINT8: return (CT)v->u.uint8_val;                                   \
            case T_UINT16: return (CT)v->u.uint16_val;                                 \
            case T_UINT32: case T_TIMESTAMP: case T_DATE: return (CT)v->u.uint32_val;  \
            case T_UINT64: case T_DATETIME: return (CT)v->u.uint64_val;                \
            case T_FLOAT: return (CT)v->u.float_val;                                   \
            case T_DOUBLE: return (CT)v->u.double_val;                                 \
            default: return (CT)0;                                                     \
        }                                                                              \
    }
GET_NUM(num_i8, int8_t) GET_NUM(num_i16, int16_t) GET_NUM(num_i32, int32_t) GET_NUM(num_i64, int64_t)
GET_NUM(num_u8, uint8_t) GET_NUM(num_u16, uint16_t) GET_NUM(num_u32, uint32_t) GET_NUM(num_u64, uint64_t)
GET_NUM(num_f32, float) GET_NUM(num_f64, double)
static int num_bool(const ExprValue* v) { /* static_cast<bool>: non-zero -> true */
    switch (v->type) {
        case T_FLOAT: return v->u.float_val != 0.0f;
        case T_DOUBLE: return v->u.double_val != 0.0;
        default: return num_u64(v) != 0;
    }
}

/* =========================== date / time encodings ===========================
 * include/common/datetime.h:28-33,56-68 (layouts) and src/common/datetime.cpp (conversions; the reference fixes the
 * zone at UTC+8 without DST: mktime_fixed_r / localtime_fixed_r default tz_offset_hours = 8).
 *   DATETIME  u64: (year*13+month)<<46 | day<<41 | hour<<36 | minute<<30 | second<<24 | microsecond
 *   DATE      u32: DATETIME >> 41          TIMESTAMP u32: seconds since the epoch
 *   TIME      i32: sign * (hour<<12 | minute<<6 | second) */
static int64_t dt_div_floor(int64_t a, int64_t b) { return (a >= 0 ? a : a + 1 - b) / b; }             /* datetime.cpp:24-26 */
static int64_t dt_days_from_civil(int64_t y, unsigned m, unsigned d) {                                 /* datetime.cpp:31-38 */
    y -= m <= 2;
    const int64_t era = dt_div_floor(y, 400);
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
static int64_t dt_mktime_fixed(int year, unsigned mon, unsigned mday, int hour, int min, int sec, int tz) { /* datetime.cpp:41-52 */
    return dt_days_from_civil(year, mon, mday) * 86400LL + hour * 3600LL + min * 60LL + sec - tz * 3600LL;
}
typedef struct { int year, mon, mday, hour, min, sec; } DtCivil;
static DtCivil dt_localtime_fixed(int64_t timep, int tz) {                                             /* datetime.cpp:54-98 */
    int64_t t = timep + tz * 3600LL, days = t / 86400LL, rem = t % 86400LL;
    if (rem < 0) { rem += 86400LL; --days; }
    const int64_t z = days + 719468, era = dt_div_floor(z, 146097);
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int year = (int)(yoe + era * 400);
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    DtCivil c;
    c.mday = (int)(doy - (153 * mp + 2) / 5 + 1);
    c.mon = (int)(mp + (mp < 10 ? 3 : -9));
    c.year = year + (c.mon <= 2);
    c.hour = (int)(rem / 3600); rem %= 3600; c.min = (int)(rem / 60); c.sec = (int)(rem % 60);
    return c;
}
static uint64_t dt_pack(uint64_t year, uint64_t month, uint64_t day, uint64_t hour, uint64_t minute, uint64_t second, uint64_t macrosec) {
    return ((year * 13 + month) << 46) | (day << 41) | (hour << 36) | (minute << 30) | (second << 24) | macrosec;
}
/* str_to_datetime_internal, datetime.cpp:149-263 (the sscanf formats are the reference's) */
static uint64_t dt_str_to_datetime(const char* str_time, size_t length, int* is_full_datetime) {
    int is_full = 0;
    while (*str_time == ' ') str_time++;
    enum { max_time_size = 26 };
    size_t len = length < max_time_size ? length : max_time_size;
    char buf[max_time_size + 1]; memset(buf, 0, sizeof buf);
    memcpy(buf, str_time, strnlen(str_time, len));
    int has_delim = 1, delim_cnt = 0;
    if (isdigit((unsigned char)buf[2]) && isdigit((unsigned char)buf[4])) has_delim = 0;
    if (buf[3] == '-') has_delim = 1;
    int32_t year_length = -1; uint32_t idx = 0;
    for (; idx < len; ++idx) {
        if (has_delim) {
            if (!isdigit((unsigned char)buf[idx])) { delim_cnt++; if (year_length == -1) year_length = (int32_t)idx; }
            if (delim_cnt > 5 && buf[idx] == '.') break;
        } else if (buf[idx] == '.') break;
    }
    if (idx < len) for (uint32_t i = idx + 1; i <= idx + 6 && i < max_time_size; ++i) if (!isdigit((unsigned char)buf[i])) buf[i] = '0';
    unsigned long year = 0, month = 0, day = 0, hour = 0, minute = 0, second = 0, macrosec = 0;
    if (has_delim) {
        sscanf(buf, "%4lu%*[^0-9a-z]%2lu%*[^0-9a-z]%2lu%*[^0-9a-z]%2lu%*[^0-9a-z]%2lu%*[^0-9a-z]%2lu.%6lu",
               &year, &month, &day, &hour, &minute, &second, &macrosec);
        is_full = 1;
    } else {
        if (idx <= 6) { sscanf(buf, "%2lu%2lu%2lu", &year, &month, &day); year_length = 2; }
        else if (idx == 8) sscanf(buf, "%4lu%2lu%2lu", &year, &month, &day);
        else if (idx == 12) { sscanf(buf, "%2lu%2lu%2lu%2lu%2lu%2lu.%6lu", &year, &month, &day, &hour, &minute, &second, &macrosec); is_full = 1; year_length = 2; }
        else if (idx <= 13) { sscanf(buf, "%2lu%2lu%2lu%2lu%2lu%2lu", &year, &month, &day, &hour, &minute, &second); is_full = 1; year_length = 2; }
        else if (idx >= 14) { sscanf(buf, "%4lu%2lu%2lu%2lu%2lu%2lu.%6lu", &year, &month, &day, &hour, &minute, &second, &macrosec); is_full = 1; }
        else return 0;
    }
    if (year_length == 2) { if (year >= 70 && year < 100) year += 1900; else if (year < 70 && year > 0) year += 2000; }
    if (month > 12 || day > 31 || hour > 23 || minute > 59 || second > 59) return 0;
    if (is_full_datetime) *is_full_datetime = is_full;
    return dt_pack(year, month, day, hour, minute, second, macrosec);
}
static int64_t dt_datetime_to_timestamp(uint64_t datetime) {                                           /* datetime.cpp:304-331 */
    if (datetime == 0) return 0;
    const int year_month = (int)((datetime >> 46) & 0x1FFFF);
    const int year = year_month / 13, mon = year_month % 13, mday = (int)((datetime >> 41) & 0x1F);
    if (mon == 0 || mday == 0) return 0;
    const int64_t t = dt_mktime_fixed(year, (unsigned)mon, (unsigned)mday, (int)((datetime >> 36) & 0x1F), (int)((datetime >> 30) & 0x3F), (int)((datetime >> 24) & 0x3F), 8);
    return t <= 0 ? 0 : t;
}
static uint64_t dt_timestamp_to_datetime(int64_t timestamp) {                                          /* datetime.cpp:352-373 */
    if (timestamp == 0) return 0;
    const DtCivil c = dt_localtime_fixed(timestamp, 8);
    return dt_pack((uint64_t)c.year, (uint64_t)c.mon, (uint64_t)c.mday, (uint64_t)c.hour, (uint64_t)c.min, (uint64_t)c.sec, 0);
}
static uint32_t dt_datetime_to_date(uint64_t datetime) { return (uint32_t)((datetime >> 41) & 0x3FFFFF); } /* datetime.h:62-64 */
static uint64_t dt_date_to_datetime(uint32_t date) { return (uint64_t)date << 41; }                         /* datetime.h:65-67 */
static int32_t dt_datetime_to_time(uint64_t datetime) {                                                /* datetime.cpp:410-419 */
    return (int32_t)(((datetime >> 24) & 0x3F) | (((datetime >> 30) & 0x3F) << 6) | (((datetime >> 36) & 0x1F) << 12));
}
static uint64_t dt_time_to_datetime(int32_t tm) {                       /* datetime.cpp:420-442 — relative to today's date */
    int64_t now = (int64_t)time(NULL);
    now = ((now + 28800) / 86400) * 86400;
    int minus = 0; if (tm < 0) { minus = 1; tm = -tm; }
    int32_t delta = (int32_t)(((tm >> 12) & 0x3FF) * 3600 + ((tm >> 6) & 0x3F) * 60 + (tm & 0x3F));
    if (minus) delta = -delta;
    return dt_timestamp_to_datetime(now - 28800 + delta);
}
static int32_t dt_str_to_time(const char* str_time, size_t length) {                                   /* datetime.cpp:477-560 */
    while (*str_time == ' ') { str_time++; length--; }
    int minus = 0;
    if (str_time[0] == '-') { minus = 1; str_time++; length--; }
    size_t len = length < 20 ? length : 20;
    int day = 0, hour = 0, minute = 0, second = 0; int32_t tm = 0;
    int has_blank = 0, has_delim = 0; uint32_t idx = 0;
    for (; idx < len; ++idx) {
        if (str_time[idx] == ' ') { has_blank = 1; has_delim = 1; }
        if (str_time[idx] == ':') has_delim = 1;
        if (str_time[idx] == '.') break;
    }
    if (idx >= 12) {
        int full = 0; uint64_t d = dt_str_to_datetime(str_time, length, &full);
        if (full) return dt_datetime_to_time(d);
    }
    if (has_blank) sscanf(str_time, "%d %u:%2u:%2u", &day, (unsigned*)&hour, (unsigned*)&minute, (unsigned*)&second);
    else if (has_delim) sscanf(str_time, "%d:%2u:%2u", &hour, (unsigned*)&minute, (unsigned*)&second);
    else {
        char t[24];
        if (idx >= 4) {
            idx -= 2; memcpy(t, str_time + idx, 2); t[2] = 0; second = (int)strtoll(t, NULL, 10);
            idx -= 2; memcpy(t, str_time + idx, 2); t[2] = 0; minute = (int)strtoll(t, NULL, 10);
            memcpy(t, str_time, idx); t[idx] = 0; hour = (int)strtoll(t, NULL, 10);
        } else if (idx >= 2) {
            idx -= 2; memcpy(t, str_time + idx, 2); t[2] = 0; second = (int)strtoll(t, NULL, 10);
            memcpy(t, str_time, idx); t[idx] = 0; minute = (int)strtoll(t, NULL, 10);
        } else { memcpy(t, str_time, idx); t[idx] = 0; second = (int)strtoll(t, NULL, 10); }
    }
    if (day < 0 || hour < 0 || minute < 0 || minute > 59 || second < 0 || second > 59) return 0;
    hour += day * 24;
    tm |= second; tm |= (minute << 6); tm |= (hour << 12);
    return minus ? -tm : tm;
}
/* formatting, only to pin the restatement against the reference's own test vectors (test/test_date_time.cpp) */
static void dt_datetime_to_str(uint64_t d, int precision_len, char out[32]) {                          /* datetime.cpp:117-139 */
    const int ym = (int)((d >> 46) & 0x1FFFF), macrosec = (int)(d & 0xFFFFFF);
    snprintf(out, 32, "%04d-%02d-%02d %02d:%02d:%02d.%06d", ym / 13, ym % 13, (int)((d >> 41) & 0x1F), (int)((d >> 36) & 0x1F), (int)((d >> 30) & 0x3F), (int)((d >> 24) & 0x3F), macrosec);
    if (precision_len > 0 && precision_len <= 6) out[20 + precision_len] = 0;
    else if (precision_len == 0 || macrosec == 0) out[19] = 0;
    else out[26] = 0;
}
static void dt_timestamp_to_str(int64_t ts, char out[32]) {                                            /* datetime.cpp:100-114 */
    if (ts <= 0) { snprintf(out, 32, "0000-00-00 00:00:00"); return; }
    const DtCivil c = dt_localtime_fixed(ts, 8);
    snprintf(out, 32, "%04d-%02d-%02d %02d:%02d:%02d", c.year, c.mon, c.mday, c.hour, c.min, c.sec);
}
static void dt_time_to_str(int32_t tm, char out[32]) {                                                 /* datetime.cpp:443-455 */
    int minus = 0; if (tm < 0) { minus = 1; tm = -tm; }
    snprintf(out, 32, "%s%02d:%02d:%02d", minus ? "-" : "", (tm >> 12) & 0x3FF, (tm >> 6) & 0x3F, tm & 0x3F);
}
/* test entry points (ctypes): op 0 str->DATETIME, 1 str->TIMESTAMP (str_to_timestamp = datetime_to_timestamp(str_to_datetime), datetime.cpp:116),
 * 2 str->DATE, 3 str->TIME; *_to_str: kind 0 DATETIME (precision in `arg`), 1 TIMESTAMP, 2 TIME */
uint64_t bk_oracle_parse_datetime(const char* s, int64_t len, int op) {
    switch (op) {
        case 0: return dt_str_to_datetime(s, (size_t)len, NULL);
        case 1: return (uint64_t)(uint32_t)dt_datetime_to_timestamp(dt_str_to_datetime(s, (size_t)len, NULL));   /* stored in uint32_val, expr_value.h:551 */
        case 2: return dt_datetime_to_date(dt_str_to_datetime(s, (size_t)len, NULL));
        default: return (uint64_t)(int64_t)dt_str_to_time(s, (size_t)len);
    }
}
void bk_oracle_datetime_to_str(uint64_t v, int kind, int arg, char* out) {
    if (kind == 0) dt_datetime_to_str(v, arg, out); else if (kind == 1) dt_timestamp_to_str((int64_t)v, out); else dt_time_to_str((int32_t)v, out);
}

static int is_dt_family(int t) { return t == T_DATETIME || t == T_TIMESTAMP || t == T_DATE || t == T_TIME; }
static ExprValue* ev_cast_to(ExprValue* v, int t);
/* the value as a DATETIME image: the `case pb::DATETIME` arm of cast_to, expr_value.h:534-548 */
static uint64_t ev_as_datetime(const ExprValue* o) {
    switch (o->type) {
        case T_STRING: return o->str ? dt_str_to_datetime(o->str, strlen(o->str), NULL) : 0;
        case T_TIMESTAMP: return dt_timestamp_to_datetime((int64_t)o->u.uint32_val);
        case T_DATE: return dt_date_to_datetime(o->u.uint32_val);
        case T_TIME: return dt_time_to_datetime(o->u.int32_val);
        case T_DATETIME: return o->u.uint64_val;
        default: return num_u64(o);
    }
}
static int ev_is_numberic(const ExprValue* v) { return is_int(v->type) || v->type == T_BOOL || is_double(v->type); } /* expr_value.h:1059-1061 */

/* cast_to, expr_value.h:502-611 (STRING sources: literals only; a STRING -> integer cast is strtoull, -> double strtod) */
static ExprValue* ev_cast_to(ExprValue* v, int t) {
    if (ev_is_null(v) || v->type == T_MAXVALUE || v->type == t) return v;
    ExprValue o = *v;
    memset(&v->u, 0, sizeof v->u);
    if (o.type == T_STRING && o.str && !is_dt_family(t) && t != T_STRING) {   /* get_numberic<T> of a STRING, expr_value.h:373-383 */
        if (is_double(t)) { o.type = T_DOUBLE; o.u.double_val = strtod(o.str, NULL); }
        else { o.type = T_UINT64; o.u.uint64_val = strtoull(o.str, NULL, 10); }
        o.str = NULL;
    }
    switch (t) {
        case T_BOOL: v->u.bool_val = (uint8_t)num_bool(&o); break;
        case T_INT8: v->u.int8_val = num_i8(&o); break;
        case T_INT16: v->u.int16_val = num_i16(&o); break;
        case T_INT32: v->u.int32_val = num_i32(&o); break;
        case T_INT64: v->u.int64_val = num_i64(&o); break;
        case T_UINT8: v->u.uint8_val = num_u8(&o); break;
        case T_UINT16: v->u.uint16_val = num_u16(&o); break;
        case T_UINT32: v->u.uint32_val = num_u32(&o); break;
        case T_UINT64: v->u.uint64_val = num_u64(&o); break;
        case T_DATETIME: v->u.uint64_val = ev_as_datetime(&o); break;
        case T_TIMESTAMP: v->u.uint32_val = ev_is_numberic(&o) ? num_u32(&o) : (uint32_t)dt_datetime_to_timestamp(ev_as_datetime(&o)); break;
        case T_DATE: v->u.uint32_val = ev_is_numberic(&o) ? num_u32(&o) : dt_datetime_to_date(ev_as_datetime(&o)); break;
        case T_TIME:
            if (ev_is_numberic(&o)) v->u.int32_val = num_i32(&o);
            else if (o.type == T_STRING) v->u.int32_val = o.str ? dt_str_to_time(o.str, strlen(o.str)) : 0;
            else v->u.int32_val = dt_datetime_to_time(ev_as_datetime(&o));
            break;
        case T_FLOAT: v->u.float_val = num_f32(&o); break;
        case T_DOUBLE: v->u.double_val = num_f64(&o); break;
        default: v->u = o.u; break;
    }
    if (t != T_STRING) v->str = NULL;
    v->type = t;
    return v;
}
/* test entry point: cast_to between two non-STRING types on the 64-bit canonical images the GPU library uses */
uint64_t bk_oracle_cast_image(uint64_t image, int from, int to) {
    ExprValue v = ev_typed(from);
    switch (from) {
        case T_BOOL: v.u.bool_val = image != 0; break;
        case T_INT8: v.u.int8_val = (int8_t)image; break;
        case T_INT16: v.u.int16_val = (int16_t)image; break;
        case T_INT32: case T_TIME: v.u.int32_val = (int32_t)image; break;
        case T_INT64: v.u.int64_val = (int64_t)image; break;
        case T_UINT8: v.u.uint8_val = (uint8_t)image; break;
        case T_UINT16: v.u.uint16_val = (uint16_t)image; break;
        case T_UINT32: case T_TIMESTAMP: case T_DATE: v.u.uint32_val = (uint32_t)image; break;
        case T_FLOAT: { double d; memcpy(&d, &image, 8); v.u.float_val = (float)d; } break;
        case T_DOUBLE: memcpy(&v.u.double_val, &image, 8); break;
        default: v.u.uint64_val = image; break;
    }
    ev_cast_to(&v, to);
    switch (to) {
        case T_BOOL: return v.u.bool_val;
        case T_INT8: return (uint64_t)(int64_t)v.u.int8_val;
        case T_INT16: return (uint64_t)(int64_t)v.u.int16_val;
        case T_INT32: case T_TIME: return (uint64_t)(int64_t)v.u.int32_val;
        case T_UINT8: return v.u.uint8_val;
        case T_UINT16: return v.u.uint16_val;
        case T_UINT32: case T_TIMESTAMP: case T_DATE: return v.u.uint32_val;
        case T_FLOAT: { double d = (double)v.u.float_val; uint64_t b; memcpy(&b, &d, 8); return b; }
        case T_DOUBLE: { uint64_t b; memcpy(&b, &v.u.double_val, 8); return b; }
        default: return v.u.uint64_val;
    }
}

/* Literal::cast_to_col_type, literal.h:204-210: a numeric literal meeting a date/time type goes through its decimal string */
static void lit_cast_to_col_type(ExprValue* v, int t, char* scratch /* >= 32 bytes, lives as long as v */) {
    if (is_dt_family(t) && ev_is_numberic(v) && !is_double(v->type)) {
        if (is_uint(v->type)) snprintf(scratch, 32, "%llu", (unsigned long long)num_u64(v));
        else snprintf(scratch, 32, "%lld", (long long)num_i64(v));           /* std::to_string, expr_value.h:709-726 */
        memset(&v->u, 0, sizeof v->u); v->type = T_STRING; v->str = scratch;
    }
    ev_cast_to(v, t);
}

/* add, expr_value.h:840-881 (the BOOL arm mutates the argument in the reference; kept) */
static void ev_add(ExprValue* a, ExprValue* b) {
    switch (a->type) {
        case T_BOOL: b->u.bool_val = (uint8_t)(b->u.bool_val + num_bool(b)); return;
        case T_INT8: a->u.int8_val = (int8_t)(a->u.int8_val + num_i8(b)); return;
        case T_INT16: a->u.int16_val = (int16_t)(a->u.int16_val + num_i16(b)); return;
        case T_INT32: a->u.int32_val = (int32_t)((uint32_t)a->u.int32_val + (uint32_t)num_i32(b)); return;
        case T_INT64: a->u.int64_val = (int64_t)((uint64_t)a->u.int64_val + (uint64_t)num_i64(b)); return;
        case T_UINT8: a->u.uint8_val = (uint8_t)(a->u.uint8_val + num_u8(b)); return;
        case T_UINT16: a->u.uint16_val = (uint16_t)(a->u.uint16_val + num_u16(b)); return;
        case T_UINT32: a->u.uint32_val += num_u32(b); return;
        case T_UINT64: a->u.uint64_val += num_u64(b); return;
        case T_FLOAT: a->u.float_val += num_f32(b); return;
        case T_DOUBLE: a->u.double_val += num_f64(b); return;
        case T_NULL: *a = *b; return;
        default: return;
    }
}

/* compare, expr_value.h:892-943: NaN compares "equal" (neither > nor <) */
static int64_t ev_compare(const ExprValue* a, const ExprValue* b) {
    if (a->type == T_MAXVALUE || b->type == T_MAXVALUE) {
        if (a->type == T_MAXVALUE && b->type == T_MAXVALUE) return 0;
        return a->type == T_MAXVALUE ? 1 : -1;
    }
    switch (a->type) {
        case T_BOOL: return (int64_t)a->u.bool_val - (int64_t)b->u.bool_val;
        case T_INT8: return (int64_t)a->u.int8_val - (int64_t)b->u.int8_val;
        case T_INT16: return (int64_t)a->u.int16_val - (int64_t)b->u.int16_val;
        case T_INT32: case T_TIME: return (int64_t)a->u.int32_val - (int64_t)b->u.int32_val;
        case T_INT64: return a->u.int64_val > b->u.int64_val ? 1 : (a->u.int64_val < b->u.int64_val ? -1 : 0);
        case T_UINT8: return (int64_t)a->u.uint8_val - (int64_t)b->u.uint8_val;
        case T_UINT16: return (int64_t)a->u.uint16_val - (int64_t)b->u.uint16_val;
        case T_UINT32: case T_TIMESTAMP: case T_DATE: return (int64_t)a->u.uint32_val - (int64_t)b->u.uint32_val;
        case T_UINT64: case T_DATETIME:
            return a->u.uint64_val > b->u.uint64_val ? 1 : (a->u.uint64_val < b->u.uint64_val ? -1 : 0);
        case T_FLOAT: return a->u.float_val > b->u.float_val ? 1 : (a->u.float_val < b->u.float_val ? -1 : 0);
        case T_DOUBLE: return a->u.double_val > b->u.double_val ? 1 : (a->u.double_val < b->u.double_val ? -1 : 0);
        case T_NULL: return b->type == T_NULL ? 0 : -1;
        default: return 0;
    }
}

/* compare_diff_type, expr_value.h:954-989 (numeric arms) */
static int64_t ev_compare_diff_type(ExprValue* a, ExprValue* b) {
    if (a->type == b->type) return ev_compare(a, b);
    if (is_int(a->type) && is_int(b->type)) {
        if (is_uint(a->type) || is_uint(b->type)) { ev_cast_to(a, T_UINT64); ev_cast_to(b, T_UINT64); }
        else { ev_cast_to(a, T_INT64); ev_cast_to(b, T_INT64); }
    } else if (a->type == T_DATETIME || b->type == T_DATETIME) { ev_cast_to(a, T_DATETIME); ev_cast_to(b, T_DATETIME);
    } else if (a->type == T_TIMESTAMP || b->type == T_TIMESTAMP) { ev_cast_to(a, T_TIMESTAMP); ev_cast_to(b, T_TIMESTAMP);
    } else if (a->type == T_DATE || b->type == T_DATE) { ev_cast_to(a, T_DATE); ev_cast_to(b, T_DATE);
    } else if (a->type == T_TIME || b->type == T_TIME) { ev_cast_to(a, T_TIME); ev_cast_to(b, T_TIME);
    } else { ev_cast_to(a, T_DOUBLE); ev_cast_to(b, T_DOUBLE); }
    return ev_compare(a, b);
}

int64_t bko_ev_compare(int ta, uint64_t ba, int tb, uint64_t bb, int diff_type) {
    ExprValue a = ev_typed(ta), b = ev_typed(tb);
    memcpy(&a.u, &ba, 8); memcpy(&b.u, &bb, 8);
    return diff_type ? ev_compare_diff_type(&a, &b) : ev_compare(&a, &b);
}
uint64_t bko_ev_cast(int from_type, uint64_t bits, int to_type) {
    ExprValue a = ev_typed(from_type); memcpy(&a.u, &bits, 8);
    ev_cast_to(&a, to_type);
    uint64_t out = 0; memcpy(&out, &a.u, 8); return out;
}

/* =========================== MutTableKey =========================== */
typedef struct { uint8_t* p; size_t n, cap; } Bytes;
static void bytes_put(Bytes* b, const void* src, size_t n) {
    if (b->n + n > b->cap) { b->cap = (b->cap + n) * 2 + 16; b->p = (uint8_t*)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, src, n); b->n += n;
}
static void put_be(Bytes* b, uint64_t v, int nbytes) {
    uint8_t t[8];
    for (int i = 0; i < nbytes; i++) t[i] = (uint8_t)(v >> (8 * (nbytes - 1 - i)));
    bytes_put(b, t, (size_t)nbytes);
}
static uint64_t encode_f64(double d) { /* key_encoder.h: sign-magnitude -> memcomparable */
    uint64_t u; memcpy(&u, &d, 8);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
static uint32_t encode_f32(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
/* append_value, mut_table_key.h:167-200: NULL appends nothing */
static void key_append_value(Bytes* b, const ExprValue* v) {
    switch (v->type) {
        case T_BOOL: { uint8_t e = v->u.bool_val ? 1 : 0; bytes_put(b, &e, 1); } break;
        case T_INT8: put_be(b, (uint8_t)v->u.int8_val ^ 0x80u, 1); break;
        case T_INT16: put_be(b, (uint16_t)v->u.int16_val ^ 0x8000u, 2); break;
        case T_INT32: case T_TIME: put_be(b, (uint32_t)v->u.int32_val ^ 0x80000000u, 4); break;
        case T_INT64: put_be(b, (uint64_t)v->u.int64_val ^ 0x8000000000000000ull, 8); break;
        case T_UINT8: put_be(b, v->u.uint8_val, 1); break;
        case T_UINT16: put_be(b, v->u.uint16_val, 2); break;
        case T_UINT32: case T_TIMESTAMP: case T_DATE: put_be(b, v->u.uint32_val, 4); break;
        case T_UINT64: case T_DATETIME: put_be(b, v->u.uint64_val, 8); break;
        case T_FLOAT: put_be(b, encode_f32(v->u.float_val), 4); break;
        case T_DOUBLE: put_be(b, encode_f64(v->u.double_val), 8); break;
        default: break;
    }
}
int bko_key_encode(int type, uint64_t bits, uint8_t out[8]) {
    ExprValue a = ev_typed(type); memcpy(&a.u, &bits, 8);
    Bytes b = {0, 0, 0}; key_append_value(&b, &a);
    int n = (int)b.n; if (n > 0) memcpy(out, b.p, (size_t)n); free(b.p); return n;
}

/* =========================== plan description =========================== */
typedef struct { int tuple_id, n_slots; int* slot_ids; int* types; } TupleDesc;

#define MAX_FN_ARGS 16
typedef struct Expr {
    int node_type, col_type, nchildren;
    struct Expr** children;
    int tuple_id, slot_id;            /* SLOT_REF / AGG_EXPR */
    ExprValue lit;                    /* literals */
    char* str_lit; char lit_scratch[32];  /* STRING literal bytes; decimal image of a numeric literal cast to a date/time type */
    int fn_op; char name[64];
    int n_arg_types, arg_types[4], return_type;
    int agg_type, final_slot, inter_slot;
    /* InPredicate */
    int map_type, has_null; int64_t* int_set; double* dbl_set; int set_n;
    int is_constant;
    int is_distinct;                  /* AGG_EXPR *_distinct: a merger updates it from its input rows (no merge), agg_fn_call.cpp:719-727 */
} Expr;

typedef struct Node {
    int node_type, nchildren; int64_t limit;
    struct Node** children;
    int tuple_id;                     /* scan / sort */
    int n_conj; Expr** conj;          /* filter / join conditions */
    int agg_tuple_id, n_group, n_agg; Expr** group; Expr** aggs; /* agg */
    int n_order; Expr** order; int* is_asc; int* is_null_first;  /* sort */
    int join_type; int64_t offset;
    void* st;                         /* runtime state */
    int64_t num_rows_returned;
} Node;

typedef struct {
    const int32_t* w; size_t n, pos; int fail; char* err; size_t errlen;
} Reader;
static int32_t rd(Reader* r) { if (r->pos >= r->n) { r->fail = 1; return 0; } return r->w[r->pos++]; }
static int64_t rd64(Reader* r) { uint32_t lo = (uint32_t)rd(r); uint32_t hi = (uint32_t)rd(r); return (int64_t)(((uint64_t)hi << 32) | lo); }
static void rdstr(Reader* r, char* out, size_t cap) {
    int32_t len = rd(r); size_t words = ((size_t)len + 3) / 4;
    if (len < 0 || r->pos + words > r->n) { r->fail = 1; out[0] = 0; return; }
    size_t c = (size_t)len < cap - 1 ? (size_t)len : cap - 1;
    memcpy(out, (const char*)(r->w + r->pos), c); out[c] = 0; r->pos += words;
}
static void set_err(Reader* r, const char* m) { if (r->err && r->errlen) snprintf(r->err, r->errlen, "%s", m); r->fail = 1; }

static Expr* parse_enode(Reader* r, int* remaining) {
    if (*remaining <= 0) { set_err(r, "expr node list too short"); return NULL; }
    (*remaining)--;
    Expr* e = (Expr*)calloc(1, sizeof(Expr));
    e->node_type = rd(r); e->col_type = rd(r); e->nchildren = rd(r);
    e->lit = ev_null();
    switch (e->node_type) {
        case E_SLOT_REF: e->tuple_id = rd(r); e->slot_id = rd(r); break;
        case E_NULL_LITERAL: e->lit = ev_null(); break;
        case E_BOOL_LITERAL: e->lit = ev_bool(rd(r)); break;
        case E_INT_LITERAL: e->lit = ev_typed(T_INT64); e->lit.u.int64_val = rd64(r); break;
        case E_DOUBLE_LITERAL: { int64_t b = rd64(r); e->lit = ev_typed(T_DOUBLE); memcpy(&e->lit.u.double_val, &b, 8); } break;
        case E_STRING_LITERAL: { /* literal.h:69-73 */
            int32_t len = r->pos < r->n ? r->w[r->pos] : -1;
            if (len < 0 || len > 4096) { set_err(r, "bad string literal"); break; }
            e->str_lit = (char*)calloc((size_t)len + 1, 1);
            rdstr(r, e->str_lit, (size_t)len + 1);
            e->lit = ev_typed(T_STRING); e->lit.str = e->str_lit;
        } break;
        case E_DATETIME_LITERAL: e->lit = ev_typed(T_DATETIME); e->lit.u.uint64_val = (uint64_t)rd64(r); break;  /* literal.h:95-114 */
        case E_TIME_LITERAL: e->lit = ev_typed(T_TIME); e->lit.u.int32_val = (int32_t)rd64(r); break;
        case E_TIMESTAMP_LITERAL: e->lit = ev_typed(T_TIMESTAMP); e->lit.u.uint32_val = (uint32_t)rd64(r); break;
        case E_DATE_LITERAL: e->lit = ev_typed(T_DATE); e->lit.u.uint32_val = (uint32_t)rd64(r); break;
        case E_AGG_EXPR:
            rdstr(r, e->name, sizeof e->name);
            e->tuple_id = rd(r); e->final_slot = rd(r); e->inter_slot = rd(r);
            if (!strcmp(e->name, "count_star")) e->agg_type = A_COUNT_STAR;
            else if (!strcmp(e->name, "count")) e->agg_type = A_COUNT;
            else if (!strcmp(e->name, "count_distinct")) { e->agg_type = A_COUNT; e->is_distinct = 1; }   /* name_type_map + _is_distinct, agg_fn_call.cpp:32-80 */
            else if (!strcmp(e->name, "sum_distinct")) { e->agg_type = A_SUM; e->is_distinct = 1; }
            else if (!strcmp(e->name, "avg_distinct")) { e->agg_type = A_AVG; e->is_distinct = 1; }
            else if (!strcmp(e->name, "sum")) e->agg_type = A_SUM;
            else if (!strcmp(e->name, "avg")) e->agg_type = A_AVG;
            else if (!strcmp(e->name, "min")) e->agg_type = A_MIN;
            else if (!strcmp(e->name, "max")) e->agg_type = A_MAX;
            else set_err(r, "unsupported aggregate");
            break;
        case E_FUNCTION_CALL: case E_IS_NULL: case E_IN: case E_NOT: case E_AND: case E_OR:
        case E_XOR: case E_IS_TRUE:
            e->fn_op = rd(r); rdstr(r, e->name, sizeof e->name);
            e->n_arg_types = rd(r);
            if (e->n_arg_types < 0 || e->n_arg_types > 4) { set_err(r, "bad n_arg_types"); e->n_arg_types = 0; }
            for (int i = 0; i < e->n_arg_types; i++) e->arg_types[i] = rd(r);
            e->return_type = rd(r);
            break;
        default: set_err(r, "unsupported expr node type"); break;
    }
    if (e->nchildren < 0 || e->nchildren > 4096) { set_err(r, "bad num_children"); e->nchildren = 0; }
    e->children = (Expr**)calloc((size_t)e->nchildren + 1, sizeof(Expr*));
    for (int i = 0; i < e->nchildren && !r->fail; i++) e->children[i] = parse_enode(r, remaining);
    return e;
}
static Expr* parse_expr(Reader* r) {
    int n = rd(r);
    Expr* e = parse_enode(r, &n);
    if (!r->fail && n != 0) set_err(r, "expr node count mismatch");
    return e;
}

static Node* parse_node(Reader* r, int* remaining) {
    if (*remaining <= 0) { set_err(r, "plan node list too short"); return NULL; }
    (*remaining)--;
    Node* n = (Node*)calloc(1, sizeof(Node));
    n->node_type = rd(r); n->nchildren = rd(r); n->limit = rd64(r);
    switch (n->node_type) {
        case N_SCAN: n->tuple_id = rd(r); (void)rd64(r); break;
        case N_WHERE_FILTER: case N_TABLE_FILTER: case N_HAVING_FILTER:
            n->n_conj = rd(r); n->conj = (Expr**)calloc((size_t)n->n_conj + 1, sizeof(Expr*));
            for (int i = 0; i < n->n_conj && !r->fail; i++) n->conj[i] = parse_expr(r);
            break;
        case N_AGG: case N_MERGE_AGG:
            n->agg_tuple_id = rd(r);
            n->n_group = rd(r); n->group = (Expr**)calloc((size_t)n->n_group + 1, sizeof(Expr*));
            for (int i = 0; i < n->n_group && !r->fail; i++) n->group[i] = parse_expr(r);
            n->n_agg = rd(r); n->aggs = (Expr**)calloc((size_t)n->n_agg + 1, sizeof(Expr*));
            for (int i = 0; i < n->n_agg && !r->fail; i++) n->aggs[i] = parse_expr(r);
            break;
        case N_SORT:
            n->tuple_id = rd(r); n->n_order = rd(r);
            n->order = (Expr**)calloc((size_t)n->n_order + 1, sizeof(Expr*));
            n->is_asc = (int*)calloc((size_t)n->n_order + 1, sizeof(int));
            n->is_null_first = (int*)calloc((size_t)n->n_order + 1, sizeof(int));
            for (int i = 0; i < n->n_order && !r->fail; i++) {
                n->order[i] = parse_expr(r); n->is_asc[i] = rd(r); n->is_null_first[i] = rd(r);
            }
            break;
        case N_JOIN:
            n->join_type = rd(r); n->n_conj = rd(r);
            n->conj = (Expr**)calloc((size_t)n->n_conj + 1, sizeof(Expr*));
            for (int i = 0; i < n->n_conj && !r->fail; i++) n->conj[i] = parse_expr(r);
            break;
        case N_LIMIT: n->offset = rd64(r); break;
        case N_PACKET: case N_SELECT_MANAGER: break;
        default: set_err(r, "unsupported plan node type"); break;
    }
    if (n->nchildren < 0 || n->nchildren > 2) { set_err(r, "bad plan num_children"); n->nchildren = 0; }
    n->children = (Node**)calloc((size_t)n->nchildren + 1, sizeof(Node*));
    for (int i = 0; i < n->nchildren && !r->fail; i++) n->children[i] = parse_node(r, remaining);
    return n;
}

/* =========================== execution context =========================== */
typedef struct {
    TupleDesc tuples[MAX_TUPLES]; int n_tuples;
    const bko_column* cols; int n_cols;
    int64_t rows_scanned, rows_filtered;
    char* err; size_t errlen;
} Ctx;

static const TupleDesc* find_tuple(const Ctx* c, int tuple_id) {
    for (int i = 0; i < c->n_tuples; i++) if (c->tuples[i].tuple_id == tuple_id) return &c->tuples[i];
    return NULL;
}
static int slot_type(const Ctx* c, int tuple_id, int slot_id) {
    const TupleDesc* t = find_tuple(c, tuple_id);
    if (!t) return T_INVALID;
    for (int i = 0; i < t->n_slots; i++) if (t->slot_ids[i] == slot_id) return t->types[i];
    return T_INVALID;
}
static const bko_column* find_col(const Ctx* c, int tuple_id, int slot_id) {
    for (int i = 0; i < c->n_cols; i++)
        if (c->cols[i].tuple_id == tuple_id && c->cols[i].slot_id == slot_id) return &c->cols[i];
    return NULL;
}

/* MemRow: one row index per scan tuple (the reference holds one protobuf message per tuple,
 * include/mem_row/mem_row.h:28-215) plus, for rows that became group accumulators, the slots
 * of the aggregate tuple. */
typedef struct MemRow {
    int64_t idx[MAX_TUPLES];     /* -1 = tuple not assigned (NULL-extended join side) */
    ExprValue* agg;              /* slots of the agg tuple, index = slot position */
    int agg_tuple_id;
    int64_t arrival;
} MemRow;
static void memrow_init(MemRow* r) { for (int i = 0; i < MAX_TUPLES; i++) r->idx[i] = -1; r->agg = NULL; r->agg_tuple_id = -1; r->arrival = 0; }

/* storage type of a slot, src/common/common.cpp:514-544 (primitive_to_other_type) */
static int storage_type(int t) {
    switch (t) {
        case T_INT8: case T_INT16: case T_INT32: case T_TIME: return T_INT32;
        case T_INT64: return T_INT64;
        case T_UINT8: case T_UINT16: case T_UINT32: case T_TIMESTAMP: case T_DATE: return T_UINT32;
        case T_UINT64: case T_DATETIME: return T_UINT64;
        case T_FLOAT: return T_FLOAT; case T_DOUBLE: return T_DOUBLE;
        case T_BOOL: case T_NULL: return T_BOOL;
        default: return T_STRING;
    }
}
static int col_is_valid(const bko_column* c, int64_t i) { return !c->validity || ((c->validity[i >> 3] >> (i & 7)) & 1); }

/* MemRow::get_value on a scan tuple: value typed by its STORAGE type (message_helper.h:198-253) */
static ExprValue column_get(const bko_column* c, int64_t i) {
    if (!col_is_valid(c, i)) return ev_null();
    ExprValue v = ev_typed(storage_type(c->prim_type));
    switch (v.type) {
        case T_INT32: v.u.int32_val = ((const int32_t*)c->values)[i]; break;
        case T_INT64: v.u.int64_val = ((const int64_t*)c->values)[i]; break;
        case T_UINT32: v.u.uint32_val = ((const uint32_t*)c->values)[i]; break;
        case T_UINT64: v.u.uint64_val = ((const uint64_t*)c->values)[i]; break;
        case T_FLOAT: v.u.float_val = ((const float*)c->values)[i]; break;
        case T_DOUBLE: v.u.double_val = ((const double*)c->values)[i]; break;
        case T_BOOL: v.u.bool_val = ((const uint8_t*)c->values)[i]; break;
        case T_STRING: /* 16-byte AVG intermediate shipped by a store (MERGE_AGG input) */
            memcpy(&v.avg, (const uint8_t*)c->values + 16 * i, 16); break;
        default: break;
    }
    return v;
}

static int agg_slot_pos(const Ctx* c, int tuple_id, int slot_id) {
    const TupleDesc* t = find_tuple(c, tuple_id);
    if (!t) return -1;
    for (int i = 0; i < t->n_slots; i++) if (t->slot_ids[i] == slot_id) return i;
    return -1;
}
/* MemRow::get_value / set_value on the aggregate tuple (message_helper.h:155-253) */
static ExprValue memrow_get(const Ctx* c, const MemRow* r, int tuple_id, int slot_id) {
    if (r->agg && r->agg_tuple_id == tuple_id) {
        int p = agg_slot_pos(c, tuple_id, slot_id);
        return p < 0 ? ev_null() : r->agg[p];
    }
    if (tuple_id < 0 || tuple_id >= MAX_TUPLES || r->idx[tuple_id] < 0) return ev_null();
    const bko_column* col = find_col(c, tuple_id, slot_id);
    if (!col) return ev_null();
    return column_get(col, r->idx[tuple_id]);
}
static void memrow_set(const Ctx* c, MemRow* r, int tuple_id, int slot_id, const ExprValue* v) {
    int p = agg_slot_pos(c, tuple_id, slot_id);
    if (p < 0 || !r->agg) return;
    if (ev_is_null(v)) { r->agg[p] = ev_null(); return; }
    int st = storage_type(slot_type(c, tuple_id, slot_id));
    if (st == T_STRING) { r->agg[p] = *v; r->agg[p].type = T_STRING; return; }
    ExprValue o = *v; ev_cast_to(&o, st); o.type = st; r->agg[p] = o;
}

/* =========================== expression engine =========================== */
static int expr_constant(const Expr* e) {
    if (e->node_type == E_SLOT_REF || e->node_type == E_AGG_EXPR) return 0;
    for (int i = 0; i < e->nchildren; i++) if (!expr_constant(e->children[i])) return 0;
    return 1;
}
static int expr_is_literal(const Expr* e) {
    return (e->node_type >= E_NULL_LITERAL && e->node_type <= E_STRING_LITERAL) || e->node_type == E_TIMESTAMP_LITERAL ||
           e->node_type == E_DATETIME_LITERAL || e->node_type == E_DATE_LITERAL || e->node_type == E_TIME_LITERAL;
}
static int all_int2(const int* t, int n) { for (int i = 0; i < n; i++) if (!is_int(t[i])) return 0; return 1; }
static int has_t(const int* t, int n, int (*f)(int)) { for (int i = 0; i < n; i++) if (f(t[i])) return 1; return 0; }
static int has_eq(const int* t, int n, int v) { for (int i = 0; i < n; i++) if (t[i] == v) return 1; return 0; }
static void complete(Expr* e, int nargs, int at, int rt) { /* fn_manager.cpp:419-464 */
    e->n_arg_types = nargs; for (int i = 0; i < nargs; i++) e->arg_types[i] = at; e->return_type = rt;
}

static ExprValue expr_value(const Ctx* c, Expr* e, const MemRow* row);

/* ScalarFnCall::type_inferer + FunctionManager::complete_fn (scalar_fn_call.cpp:40-120,
 * fn_manager.cpp:316-409) and AggFnCall::type_inferer (agg_fn_call.cpp:87-122) */
static int type_infer(Ctx* c, Expr* e) {
    for (int i = 0; i < e->nchildren; i++) if (type_infer(c, e->children[i]) < 0) return -1;
    e->is_constant = expr_constant(e);
    switch (e->node_type) {
        case E_SLOT_REF:
            if (e->col_type == T_INVALID) e->col_type = slot_type(c, e->tuple_id, e->slot_id);
            if (e->col_type == T_INVALID) { snprintf(c->err, c->errlen, "unknown slot %d_%d", e->tuple_id, e->slot_id); return -1; }
            return 0;
        case E_NULL_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_NULL; return 0;
        case E_BOOL_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_BOOL; return 0;
        case E_INT_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_INT64; return 0;
        case E_DOUBLE_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_DOUBLE; return 0;
        case E_STRING_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_STRING; return 0;
        case E_DATETIME_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_DATETIME; return 0;
        case E_TIMESTAMP_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_TIMESTAMP; return 0;
        case E_DATE_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_DATE; return 0;
        case E_TIME_LITERAL: if (e->col_type == T_INVALID) e->col_type = T_TIME; return 0;
        case E_AGG_EXPR: {
            int ct = e->nchildren ? e->children[0]->col_type : T_INVALID;
            switch (e->agg_type) {
                case A_COUNT_STAR: case A_COUNT: e->col_type = T_INT64; break;
                case A_AVG: e->col_type = T_DOUBLE; break;
                case A_SUM: e->col_type = (is_double(ct) || is_string(ct)) ? T_DOUBLE : (is_uint(ct) ? T_UINT64 : T_INT64); break;
                default: e->col_type = ct; break;
            }
            return 0;
        }
        default: break;
    }
    if (e->node_type == E_IN) {
        /* InPredicate::singel_open, predicate.cpp:102-148 */
        if (e->nchildren < 2) { snprintf(c->err, c->errlen, "IN needs a list"); return -1; }
        if (e->children[0]->node_type == E_SLOT_REF)
            for (int i = 1; i < e->nchildren; i++) if (e->children[i]->is_constant) e->children[i]->col_type = e->children[0]->col_type;
        int types[2] = { e->children[0]->col_type, e->children[1]->col_type };
        if (all_int2(types, 2)) e->map_type = T_INT64;
        else if (has_eq(types, 2, T_DATETIME)) e->map_type = T_DATETIME;
        else if (has_eq(types, 2, T_TIMESTAMP)) e->map_type = T_TIMESTAMP;
        else if (has_eq(types, 2, T_DATE)) e->map_type = T_DATE;
        else if (has_eq(types, 2, T_TIME)) e->map_type = T_TIME;
        else if (has_t(types, 2, is_double) || has_t(types, 2, is_int)) e->map_type = T_DOUBLE;
        else { snprintf(c->err, c->errlen, "string IN outside the path"); return -1; }
        e->int_set = (int64_t*)calloc((size_t)e->nchildren, sizeof(int64_t));
        e->dbl_set = (double*)calloc((size_t)e->nchildren, sizeof(double));
        e->set_n = 0; e->has_null = 0;
        for (int i = 1; i < e->nchildren; i++) {
            if (!e->children[i]->is_constant) { snprintf(c->err, c->errlen, "only support in const"); return -1; }
            ExprValue v = expr_value(c, e->children[i], NULL);
            if (ev_is_null(&v)) { e->has_null = 1; continue; }
            ev_cast_to(&v, e->map_type);
            if (e->map_type != T_DOUBLE) e->int_set[e->set_n++] = num_i64(&v); else e->dbl_set[e->set_n++] = num_f64(&v);   /* predicate.cpp:128-136 */
        }
        if (e->col_type == T_INVALID) e->col_type = T_BOOL;
        return 0;
    }
    if (e->node_type == E_AND || e->node_type == E_OR || e->node_type == E_XOR || e->node_type == E_NOT ||
        e->node_type == E_IS_NULL || e->node_type == E_IS_TRUE) {
        if (e->col_type == T_INVALID) e->col_type = T_BOOL;
        return 0;
    }
    /* FUNCTION_CALL */
    if (e->n_arg_types > 0 && e->return_type != T_INVALID) { /* already completed by the db */
        if (e->col_type == T_INVALID) e->col_type = e->return_type;
        return 0;
    }
    switch (e->fn_op) { /* predicate handled as the column's type, scalar_fn_call.cpp:57-67 */
        case FT_EQ: case FT_NE: case FT_GE: case FT_GT: case FT_LE: case FT_LT:
            if (e->nchildren == 2 && e->children[0]->node_type == E_SLOT_REF && e->children[1]->is_constant)
                e->children[1]->col_type = e->children[0]->col_type;
            break;
        default: break;
    }
    int types[4]; int n = e->nchildren < 4 ? e->nchildren : 4;
    for (int i = 0; i < n; i++) types[i] = e->children[i]->col_type;
    switch (e->fn_op) {
        case FT_EQ: case FT_NE: case FT_GE: case FT_GT: case FT_LE: case FT_LT:
            if (all_int2(types, n)) complete(e, 2, has_t(types, n, is_uint) ? T_UINT64 : T_INT64, T_BOOL);
            else if (has_eq(types, n, T_DATETIME)) complete(e, 2, T_DATETIME, T_BOOL);
            else if (has_eq(types, n, T_TIMESTAMP)) complete(e, 2, T_TIMESTAMP, T_BOOL);
            else if (has_eq(types, n, T_DATE)) complete(e, 2, T_DATE, T_BOOL);
            else if (has_eq(types, n, T_TIME)) complete(e, 2, T_TIME, T_BOOL);
            else if (has_t(types, n, is_double)) complete(e, 2, T_DOUBLE, T_BOOL);
            else if (has_t(types, n, is_int)) complete(e, 2, T_DOUBLE, T_BOOL);
            else complete(e, 2, T_STRING, T_BOOL);
            break;
        case FT_ADD: case FT_MINUS: case FT_MULTIPLIES:
            if (has_t(types, n, is_double)) complete(e, 2, T_DOUBLE, T_DOUBLE);
            else if (has_t(types, n, is_uint)) complete(e, 2, T_UINT64, T_UINT64);
            else complete(e, 2, T_INT64, T_INT64);
            break;
        case FT_DIVIDES: complete(e, 2, T_DOUBLE, T_DOUBLE); break;
        case FT_MOD:
            if (has_t(types, n, is_uint)) complete(e, 2, T_UINT64, T_UINT64); else complete(e, 2, T_INT64, T_INT64);
            break;
        case FT_BIT_AND: case FT_BIT_OR: case FT_BIT_XOR: case FT_LS: case FT_RS: complete(e, 2, T_UINT64, T_UINT64); break;
        case FT_BIT_NOT: complete(e, 1, T_UINT64, T_UINT64); break;
        case FT_UMINUS:
            if (has_t(types, n, is_double)) complete(e, 1, T_DOUBLE, T_DOUBLE);
            else if (has_t(types, n, is_uint)) complete(e, 1, T_UINT64, T_UINT64);
            else complete(e, 1, T_INT64, T_INT64);
            break;
        case FT_LOGIC_NOT: complete(e, 1, T_BOOL, T_BOOL); break;
        case FT_COMMON: { /* return_type_map[fn.name()] + complete_common_fn, fn_manager.cpp:398-401,466-514; arg_types stay empty */
            int merge[MAX_FN_ARGS], nm = 0, nn = e->nchildren;
            if (nn > MAX_FN_ARGS) { snprintf(c->err, c->errlen, "%s: too many arguments", e->name); return -1; }
            if (!strcmp(e->name, "if")) { if (nn != 3) { snprintf(c->err, c->errlen, "if() needs 3 arguments"); return -1; } merge[nm++] = e->children[1]->col_type; merge[nm++] = e->children[2]->col_type; }
            else if (!strcmp(e->name, "ifnull")) { if (nn != 2) { snprintf(c->err, c->errlen, "ifnull() needs 2 arguments"); return -1; } merge[nm++] = e->children[0]->col_type; merge[nm++] = e->children[1]->col_type; }
            else if (!strcmp(e->name, "case_when")) { for (int i = 1; i < nn; i++) if (i % 2 == 1 || i + 1 == nn) merge[nm++] = e->children[i]->col_type; }
            else if (!strcmp(e->name, "abs") || !strcmp(e->name, "round") || !strcmp(e->name, "cast_to_double")) e->return_type = T_DOUBLE;
            else if (!strcmp(e->name, "floor") || !strcmp(e->name, "ceil") || !strcmp(e->name, "ceiling") || !strcmp(e->name, "cast_to_signed") ||
                     !strcmp(e->name, "sign") || !strcmp(e->name, "bit_count")) e->return_type = T_INT64;
            else if (!strcmp(e->name, "sqrt") || !strcmp(e->name, "mod") || !strcmp(e->name, "sin") || !strcmp(e->name, "asin") || !strcmp(e->name, "cos") ||
                     !strcmp(e->name, "acos") || !strcmp(e->name, "tan") || !strcmp(e->name, "cot") || !strcmp(e->name, "atan") || !strcmp(e->name, "ln") ||
                     !strcmp(e->name, "log") || !strcmp(e->name, "pi") || !strcmp(e->name, "pow") || !strcmp(e->name, "power") || !strcmp(e->name, "greatest") ||
                     !strcmp(e->name, "least")) e->return_type = T_DOUBLE;   /* return_type_map, fn_manager.cpp:105-128 */
            else if (!strcmp(e->name, "cast_to_unsigned")) e->return_type = T_UINT64;
            else { snprintf(c->err, c->errlen, "unsupported function %s", e->name); return -1; }
            if (nm) { /* has_merged_type, include/common/type_utils.h:502-560 */
                int all_null = 1, all_equal = 1, all_num = 1, has_dbl = 0, has_u64 = 0, has_sgn = 0, first = T_NULL;
                for (int i = 0; i < nm; i++) {
                    int t = merge[i];
                    if (t == T_NULL) continue;
                    if (all_null) { first = t; all_null = 0; }
                    if (t != first) all_equal = 0;
                    if (!(is_double(t) || is_int(t) || t == T_BOOL)) all_num = 0;
                    if (is_double(t)) has_dbl = 1;
                    if (t == T_UINT64) has_u64 = 1;
                    if (is_signed(t)) has_sgn = 1;
                }
                if (all_null) e->return_type = T_NULL;
                else if (all_equal) e->return_type = first;
                else if (all_num) e->return_type = has_dbl ? T_DOUBLE : (has_u64 ? (has_sgn ? T_DOUBLE : T_UINT64) : T_INT64);
                else { snprintf(c->err, c->errlen, "%s: date/time or string branches are outside the path", e->name); return -1; }
            }
            if (e->col_type == T_INVALID) e->col_type = e->return_type;
            return 0;
        }
        default: snprintf(c->err, c->errlen, "unsupported fn_op %d", e->fn_op); return -1;
    }
    if (e->col_type == T_INVALID) e->col_type = e->return_type;
    /* Literal type cast, scalar_fn_call.cpp:113-117 + Literal::cast_to_col_type */
    for (int i = 0; i < e->n_arg_types && i < e->nchildren; i++)
        if (expr_is_literal(e->children[i])) {
            Expr* l = e->children[i];
            lit_cast_to_col_type(&l->lit, e->arg_types[i], l->lit_scratch);
            if (!ev_is_null(&l->lit)) l->col_type = l->lit.type;   /* value_to_node_type: _col_type = _value.type */
        }
    return 0;
}

static int in_int(const Expr* e, int64_t v) { for (int i = 0; i < e->set_n; i++) if (e->int_set[i] == v) return 1; return 0; }
static int in_dbl(const Expr* e, double v) { for (int i = 0; i < e->set_n; i++) if (e->dbl_set[i] == v) return 1; return 0; }

/* named builtins, src/expr/internal_functions.cpp: round :52-68, floor :70-77, ceil :79-86, abs :88-99, case_when :2351-2366,
 * if_ :2383-2388, ifnull :2390-2395, cast_to_signed/unsigned/double :2941-2963 */
static ExprValue call_common(const Expr* e, ExprValue* a, int n) {
    ExprValue r;
    if (!strcmp(e->name, "if")) return (!ev_is_null(&a[0]) && num_bool(&a[0])) ? a[1] : a[2]; /* get_numberic<bool>() of NULL is false */
    if (!strcmp(e->name, "ifnull")) return ev_is_null(&a[0]) ? a[1] : a[0];
    if (!strcmp(e->name, "case_when")) {
        for (int i = 0; i < n / 2; i++) if (!ev_is_null(&a[2 * i]) && num_bool(&a[2 * i])) return a[2 * i + 1];
        return n % 2 == 0 ? ev_null() : a[n - 1];
    }
    if (!strcmp(e->name, "pi")) { r = ev_typed(T_DOUBLE); r.u.double_val = M_PI; return r; }                         /* :259-263 */
    if (!strcmp(e->name, "greatest") || !strcmp(e->name, "least")) {                                                 /* :265-317 */
        const int gt = e->name[0] == 'g'; double ret = 0; int found = 0;
        for (int i = 0; i < n; i++) {
            if (ev_is_null(&a[i])) return ev_null();
            const double v = num_f64(&a[i]);
            if (!found) { found = 1; ret = v; } else if (gt ? v > ret : v < ret) ret = v;
        }
        if (!found) return ev_null();
        r = ev_typed(T_DOUBLE); r.u.double_val = ret; return r;
    }
    if (!strcmp(e->name, "mod") || !strcmp(e->name, "log") || !strcmp(e->name, "pow") || !strcmp(e->name, "power")) {
        if (n < 2 || ev_is_null(&a[0]) || ev_is_null(&a[1])) return ev_null();
        const double p = num_f64(&a[0]), q = num_f64(&a[1]);
        r = ev_typed(T_DOUBLE);
        if (e->name[0] == 'm') { if (fabs(q - 0) < 1e-9) return ev_null(); r.u.double_val = fmod(p, q); }            /* mod :114-127, float_equal(rhs, 0) */
        else if (e->name[0] == 'l') { if (p <= 0 || q <= 0 || p == 1) return ev_null(); r.u.double_val = log(q) / log(p); } /* log(base, val) :234-246 */
        else r.u.double_val = pow(p, q);                                                                             /* :248-257 */
        return r;
    }
    if (!strcmp(e->name, "bit_count")) {                                                                             /* :336-350 */
        if (n != 1 || ev_is_null(&a[0])) return ev_null();
        ExprValue t = a[0]; ev_cast_to(&t, T_UINT64);
        r = ev_typed(T_INT64);
        for (uint64_t v = t.u.uint64_val; v; v >>= 1) r.u.int64_val += (int64_t)(v & 1);
        return r;
    }
    if (n < 1 || ev_is_null(&a[0])) return ev_null();
    {   /* one-argument DOUBLE functions :101-232: sqrt, sign, sin, asin, cos, acos, tan, cot, atan, ln */
        const double v = num_f64(&a[0]);
        r = ev_typed(T_DOUBLE);
        if (!strcmp(e->name, "sqrt")) { if (v < 0) return ev_null(); r.u.double_val = sqrt(v); return r; }
        if (!strcmp(e->name, "sign")) { r = ev_typed(T_INT64); r.u.int64_val = v > 0 ? 1 : (v < 0 ? -1 : 0); return r; }
        if (!strcmp(e->name, "sin")) { r.u.double_val = sin(v); return r; }
        if (!strcmp(e->name, "cos")) { r.u.double_val = cos(v); return r; }
        if (!strcmp(e->name, "tan")) { r.u.double_val = tan(v); return r; }
        if (!strcmp(e->name, "atan")) { r.u.double_val = atan(v); return r; }
        if (!strcmp(e->name, "asin")) { if (v < -1 || v > 1) return ev_null(); r.u.double_val = asin(v); return r; }
        if (!strcmp(e->name, "acos")) { if (v < -1 || v > 1) return ev_null(); r.u.double_val = acos(v); return r; }
        if (!strcmp(e->name, "cot")) { const double s = sin(v), c = cos(v); if (fabs(s - 0) < 1e-9) return ev_null(); r.u.double_val = c / s; return r; }
        if (!strcmp(e->name, "ln")) { if (v <= 0) return ev_null(); r.u.double_val = log(v); return r; }
    }
    if (!strcmp(e->name, "cast_to_signed")) { r = a[0]; ev_cast_to(&r, T_INT64); return r; }
    if (!strcmp(e->name, "cast_to_unsigned")) { r = a[0]; ev_cast_to(&r, T_UINT64); return r; }
    if (!strcmp(e->name, "cast_to_double")) { r = a[0]; ev_cast_to(&r, T_DOUBLE); return r; }
    double x = num_f64(&a[0]);
    if (!strcmp(e->name, "abs")) { r = ev_typed(T_DOUBLE); r.u.double_val = x < 0 ? -x : x; return r; }
    if (!strcmp(e->name, "floor")) { r = ev_typed(T_INT64); r.u.int64_val = (int64_t)floor(x); return r; }
    if (!strcmp(e->name, "ceil") || !strcmp(e->name, "ceiling")) { r = ev_typed(T_INT64); r.u.int64_val = (int64_t)ceil(x); return r; }
    if (!strcmp(e->name, "round")) {
        int bits = n == 2 ? num_i32(&a[1]) : 0;
        double base = pow(10, bits);
        r = ev_typed(T_DOUBLE);
        if (base > 0) r.u.double_val = x < 0 ? -round(-x * base) / base : round(x * base) / base;
        return r;
    }
    return ev_null();
}

static ExprValue call_fn(const Expr* e, ExprValue* a) { /* operators.cpp:18-103 */
    int at = e->arg_types[0];
    if (e->n_arg_types >= 1 && ev_is_null(&a[0])) return ev_null();
    if (e->n_arg_types >= 2 && ev_is_null(&a[1])) return ev_null();
    ExprValue r;
    switch (e->fn_op) {
        case FT_EQ: case FT_NE: case FT_GT: case FT_GE: case FT_LT: case FT_LE: {
            int lt, eq;
            if (at == T_INT64) { lt = a[0].u.int64_val < a[1].u.int64_val; eq = a[0].u.int64_val == a[1].u.int64_val; }
            else if (at == T_UINT64 || at == T_DATETIME) { lt = a[0].u.uint64_val < a[1].u.uint64_val; eq = a[0].u.uint64_val == a[1].u.uint64_val; }
            else if (at == T_DOUBLE) { /* IEEE: every ordered compare with NaN is false, != is true */
                double x = a[0].u.double_val, y = a[1].u.double_val;
                switch (e->fn_op) {
                    case FT_EQ: return ev_bool(x == y); case FT_NE: return ev_bool(x != y);
                    case FT_GT: return ev_bool(x > y); case FT_GE: return ev_bool(x >= y);
                    case FT_LT: return ev_bool(x < y); default: return ev_bool(x <= y);
                }
            } else if (at == T_TIME) { lt = a[0].u.int32_val < a[1].u.int32_val; eq = a[0].u.int32_val == a[1].u.int32_val; }
            else { lt = a[0].u.uint32_val < a[1].u.uint32_val; eq = a[0].u.uint32_val == a[1].u.uint32_val; }
            switch (e->fn_op) {
                case FT_EQ: return ev_bool(eq); case FT_NE: return ev_bool(!eq);
                case FT_GT: return ev_bool(!lt && !eq); case FT_GE: return ev_bool(!lt);
                case FT_LT: return ev_bool(lt); default: return ev_bool(lt || eq);
            }
        }
        case FT_ADD: case FT_MINUS: case FT_MULTIPLIES:
            r = ev_typed(at);
            if (at == T_DOUBLE) {
                double x = a[0].u.double_val, y = a[1].u.double_val;
                r.u.double_val = e->fn_op == FT_ADD ? x + y : (e->fn_op == FT_MINUS ? x - y : x * y);
            } else { /* int64/uint64: two's-complement wraparound */
                uint64_t x = a[0].u.uint64_val, y = a[1].u.uint64_val;
                r.u.uint64_val = e->fn_op == FT_ADD ? x + y : (e->fn_op == FT_MINUS ? x - y : x * y);
            }
            return r;
        case FT_DIVIDES: /* always DOUBLE; NULL on zero divisor */
            if (a[1].u.double_val == 0) return ev_null();
            r = ev_typed(T_DOUBLE); r.u.double_val = a[0].u.double_val / a[1].u.double_val; return r;
        case FT_MOD:
            r = ev_typed(at);
            if (at == T_UINT64) { if (a[1].u.uint64_val == 0) return ev_null(); r.u.uint64_val = a[0].u.uint64_val % a[1].u.uint64_val; }
            else {
                if (a[1].u.int64_val == 0) return ev_null();
                r.u.int64_val = a[1].u.int64_val == -1 ? 0 : a[0].u.int64_val % a[1].u.int64_val;
            }
            return r;
        case FT_BIT_AND: r = ev_typed(T_UINT64); r.u.uint64_val = a[0].u.uint64_val & a[1].u.uint64_val; return r;
        case FT_BIT_OR: r = ev_typed(T_UINT64); r.u.uint64_val = a[0].u.uint64_val | a[1].u.uint64_val; return r;
        case FT_BIT_XOR: r = ev_typed(T_UINT64); r.u.uint64_val = a[0].u.uint64_val ^ a[1].u.uint64_val; return r;
        case FT_LS: r = ev_typed(T_UINT64); r.u.uint64_val = a[1].u.uint64_val >= 64 ? 0 : a[0].u.uint64_val << a[1].u.uint64_val; return r;
        case FT_RS: r = ev_typed(T_UINT64); r.u.uint64_val = a[1].u.uint64_val >= 64 ? 0 : a[0].u.uint64_val >> a[1].u.uint64_val; return r;
        case FT_BIT_NOT: r = ev_typed(T_UINT64); r.u.uint64_val = ~a[0].u.uint64_val; return r;
        case FT_UMINUS:
            if (at == T_DOUBLE) { r = ev_typed(T_DOUBLE); r.u.double_val = -a[0].u.double_val; }
            else { r = ev_typed(T_INT64); r.u.int64_val = (int64_t)(0 - a[0].u.uint64_val); } /* minus_uint returns INT64 too */
            return r;
        case FT_LOGIC_NOT: return ev_bool(!a[0].u.bool_val);
        default: return ev_null();
    }
}

static ExprValue expr_value(const Ctx* c, Expr* e, const MemRow* row) {
    switch (e->node_type) {
        case E_SLOT_REF: { /* slot_ref.h:31-40 */
            if (!row) return ev_null();
            ExprValue v = memrow_get(c, row, e->tuple_id, e->slot_id);
            if (v.type == T_STRING) return v; /* AVG blob passes through untouched */
            ev_cast_to(&v, e->col_type); return v;
        }
        case E_NULL_LITERAL: return ev_null();
        case E_BOOL_LITERAL: case E_INT_LITERAL: case E_DOUBLE_LITERAL: case E_STRING_LITERAL:
        case E_DATETIME_LITERAL: case E_TIMESTAMP_LITERAL: case E_DATE_LITERAL: case E_TIME_LITERAL: { /* literal.h:204-206 */
            ExprValue v = e->lit; ev_cast_to(&v, e->col_type); return v;
        }
        case E_AND: { /* predicate.h:25-45 */
            int has_null = 0;
            for (int i = 0; i < e->nchildren; i++) {
                ExprValue v = expr_value(c, e->children[i], row);
                if (!ev_is_null(&v) && !num_bool(&v)) return ev_bool(0);
                if (ev_is_null(&v)) has_null = 1;
            }
            return has_null ? ev_null() : ev_bool(1);
        }
        case E_OR: { /* predicate.h:81-100 */
            int has_null = 0;
            for (int i = 0; i < e->nchildren; i++) {
                ExprValue v = expr_value(c, e->children[i], row);
                if (!ev_is_null(&v) && num_bool(&v)) return ev_bool(1);
                if (ev_is_null(&v)) has_null = 1;
            }
            return has_null ? ev_null() : ev_bool(0);
        }
        case E_XOR: {
            ExprValue a = expr_value(c, e->children[0], row), b = expr_value(c, e->children[1], row);
            if (ev_is_null(&a) || ev_is_null(&b)) return ev_null();
            return ev_bool(num_bool(&a) != num_bool(&b));
        }
        case E_NOT: { /* NotPredicate: NULL -> NULL */
            ExprValue a = expr_value(c, e->children[0], row);
            if (ev_is_null(&a)) return ev_null();
            return ev_bool(!num_bool(&a));
        }
        case E_IS_NULL: { ExprValue a = expr_value(c, e->children[0], row); return ev_bool(ev_is_null(&a)); }
        case E_IS_TRUE: { ExprValue a = expr_value(c, e->children[0], row); return ev_bool(!ev_is_null(&a) && num_bool(&a)); }
        case E_IN: { /* predicate.cpp:150-189 */
            ExprValue v = expr_value(c, e->children[0], row);
            if (ev_is_null(&v)) return ev_null();
            ev_cast_to(&v, e->map_type);
            if (e->map_type != T_DOUBLE ? in_int(e, num_i64(&v)) : in_dbl(e, num_f64(&v))) return ev_bool(1);
            return e->has_null ? ev_null() : ev_bool(0);
        }
        case E_FUNCTION_CALL: { /* scalar_fn_call.cpp:194-225 */
            ExprValue args[MAX_FN_ARGS];
            int n = e->nchildren < MAX_FN_ARGS ? e->nchildren : MAX_FN_ARGS;
            for (int i = 0; i < n; i++) args[i] = expr_value(c, e->children[i], row);
            for (int i = 0; i < e->n_arg_types && i < n; i++) ev_cast_to(&args[i], e->arg_types[i]);
            ExprValue r = e->fn_op == FT_COMMON ? call_common(e, args, n) : call_fn(e, args);
            ev_cast_to(&r, e->col_type);
            return r;
        }
        case E_AGG_EXPR: /* read as the final slot of the aggregate tuple */
            return row ? memrow_get(c, row, e->tuple_id, e->final_slot) : ev_null();
        default: return ev_null();
    }
}

/* =========================== row batches & operators =========================== */
typedef struct { MemRow* rows; int n, cap; } RowBatch;
static void batch_init(RowBatch* b, int cap) { b->rows = (MemRow*)malloc(sizeof(MemRow) * (size_t)cap); b->n = 0; b->cap = cap; }
static void batch_free(RowBatch* b) { free(b->rows); b->rows = NULL; }
static int node_get_next(Ctx* c, Node* n, RowBatch* out, int* eos);
static int node_open(Ctx* c, Node* n);
static int reached_limit(const Node* n) { return n->limit != -1 && n->num_rows_returned >= n->limit; }

/* ---- scan: synthetic column source (MockScanNode pattern, test/test_window.cpp:117-125) ---- */
typedef struct { int64_t pos, nrows; int all_tuples; } ScanState;
static int g_scan_nodes = 0;
static int scan_open(Ctx* c, Node* n) {
    ScanState* s = (ScanState*)calloc(1, sizeof *s); n->st = s;
    s->nrows = -1;
    for (int i = 0; i < c->n_cols; i++) if (c->cols[i].tuple_id == n->tuple_id) {
        if (s->nrows >= 0 && s->nrows != c->cols[i].length) { snprintf(c->err, c->errlen, "ragged columns in tuple %d", n->tuple_id); return -1; }
        s->nrows = c->cols[i].length;
    }
    /* a scalar merger's input rows carry only the aggregate tuple: the batch length is then any column's */
    if (s->nrows < 0 && g_scan_nodes == 1 && c->n_cols > 0) s->nrows = c->cols[0].length;
    if (s->nrows < 0) s->nrows = 0;
    /* A store returns rows that carry the scan tuple AND the aggregate tuple (region.cpp:3166-3216);
     * with a single scan node every input tuple therefore shares the row index (MERGE_AGG input). */
    s->all_tuples = g_scan_nodes == 1;
    return 0;
}
static int scan_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) {
    ScanState* s = (ScanState*)n->st;
    while (out->n < out->cap && s->pos < s->nrows && !reached_limit(n)) {
        MemRow* r = &out->rows[out->n++]; memrow_init(r);
        if (s->all_tuples) for (int k = 0; k < c->n_cols; k++) { int t = c->cols[k].tuple_id; if (t >= 0 && t < MAX_TUPLES && c->cols[k].length == s->nrows) r->idx[t] = s->pos; }
        r->idx[n->tuple_id] = s->pos++; n->num_rows_returned++; c->rows_scanned++;
    }
    *eos = (s->pos >= s->nrows) || reached_limit(n);
    return 0;
}

/* ---- filter ---- */
typedef struct { RowBatch child; int child_pos, child_eos; } FilterState;
static int need_copy(const Ctx* c, Node* n, const MemRow* row) { /* filter_node.cpp:726-734 */
    for (int i = 0; i < n->n_conj; i++) {
        ExprValue v = expr_value(c, n->conj[i], row);
        if (ev_is_null(&v) || !num_bool(&v)) return 0;
    }
    return 1;
}
static int filter_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) { /* filter_node.cpp:736-795 */
    FilterState* s = (FilterState*)n->st;
    for (;;) {
        if (out->n >= out->cap) return 0;
        if (s->child_pos >= s->child.n) {
            if (s->child_eos) { *eos = 1; return 0; }
            s->child.n = 0; s->child_pos = 0;
            if (node_get_next(c, n->children[0], &s->child, &s->child_eos) < 0) return -1;
            continue;
        }
        MemRow* row = &s->child.rows[s->child_pos];
        if (need_copy(c, n, row)) { out->rows[out->n++] = *row; n->num_rows_returned++; }
        else c->rows_filtered++;
        if (reached_limit(n)) { *eos = 1; return 0; }
        s->child_pos++;
    }
}

/* ---- aggregate ---- */
typedef struct { uint8_t* key; size_t klen; MemRow row; } AggEntry;
typedef struct {
    AggEntry* entries; size_t n, cap;      /* insertion order */
    int64_t* buckets; size_t nb;           /* open addressing over entry indices */
    size_t iter; int is_merger; int n_slots;
} AggState;
static uint64_t hash_bytes(const uint8_t* p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } return h ^ (h >> 29); }
static void agg_rehash(AggState* s) {
    size_t nb = s->nb ? s->nb * 2 : 256;
    int64_t* b = (int64_t*)malloc(nb * sizeof(int64_t));
    for (size_t i = 0; i < nb; i++) b[i] = -1;
    for (size_t i = 0; i < s->n; i++) {
        size_t h = hash_bytes(s->entries[i].key, s->entries[i].klen) & (nb - 1);
        while (b[h] >= 0) h = (h + 1) & (nb - 1);
        b[h] = (int64_t)i;
    }
    free(s->buckets); s->buckets = b; s->nb = nb;
}
static AggEntry* agg_seek(AggState* s, const uint8_t* key, size_t klen) {
    if (!s->nb) return NULL;
    size_t h = hash_bytes(key, klen) & (s->nb - 1);
    while (s->buckets[h] >= 0) {
        AggEntry* e = &s->entries[s->buckets[h]];
        if (e->klen == klen && !memcmp(e->key, key, klen)) return e;
        h = (h + 1) & (s->nb - 1);
    }
    return NULL;
}
static AggEntry* agg_insert(AggState* s, const uint8_t* key, size_t klen, const MemRow* row) {
    if ((s->n + 1) * 2 > s->nb) agg_rehash(s);
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 256; s->entries = (AggEntry*)realloc(s->entries, s->cap * sizeof(AggEntry)); }
    AggEntry* e = &s->entries[s->n];
    e->key = (uint8_t*)malloc(klen ? klen : 1); memcpy(e->key, key, klen); e->klen = klen; e->row = *row;
    size_t h = hash_bytes(key, klen) & (s->nb - 1);
    while (s->buckets[h] >= 0) h = (h + 1) & (s->nb - 1);
    s->buckets[h] = (int64_t)s->n; s->n++;
    return e;
}
static void encode_exprs_key(const Ctx* c, Expr** exprs, int n, const MemRow* row, Bytes* key) { /* exec_node.cpp:555-571 */
    uint8_t null_flag = 0; key->n = 0; bytes_put(key, &null_flag, 1);
    for (int i = 0; i < n; i++) {
        ExprValue v = expr_value(c, exprs[i], row);
        if (ev_is_null(&v)) { null_flag |= (uint8_t)(0x01 << (7 - i)); continue; }
        key_append_value(key, &v);
    }
    key->p[0] = null_flag;
}
static void agg_initialize(const Ctx* c, Expr* a, MemRow* dst, int only_count) { /* agg_fn_call.cpp:370-410 */
    ExprValue cur = memrow_get(c, dst, a->tuple_id, a->inter_slot);
    if (!ev_is_null(&cur)) return;
    if (a->agg_type == A_COUNT_STAR || a->agg_type == A_COUNT) { ExprValue z = ev_typed(T_INT64); memrow_set(c, dst, a->tuple_id, a->inter_slot, &z); return; }
    if (only_count) return;
    if (a->agg_type == A_AVG) {
        ExprValue z = ev_typed(T_STRING); z.avg.sum = 0; z.avg.count = 0;
        memrow_set(c, dst, a->tuple_id, a->inter_slot, &z);
        ExprValue nul = ev_null(); memrow_set(c, dst, a->tuple_id, a->final_slot, &nul);
    }
}
static void agg_update(const Ctx* c, Expr* a, const MemRow* src, MemRow* dst) { /* agg_fn_call.cpp:496-555 */
    switch (a->agg_type) {
        case A_COUNT_STAR: { ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); r.u.int64_val++; memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); return; }
        case A_COUNT: {
            for (int i = 0; i < a->nchildren; i++) { ExprValue v = expr_value(c, a->children[i], src); if (ev_is_null(&v)) return; }
            ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); r.u.int64_val++; memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); return;
        }
        case A_SUM: {
            ExprValue v = expr_value(c, a->children[0], src);
            if (!ev_is_null(&v)) { ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); ev_add(&r, &v); memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); }
            return;
        }
        case A_AVG: {
            ExprValue v = expr_value(c, a->children[0], src);
            if (!ev_is_null(&v)) { ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); r.avg.sum += num_f64(&v); r.avg.count++; memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); }
            return;
        }
        case A_MIN: case A_MAX: {
            ExprValue v = expr_value(c, a->children[0], src); ev_cast_to(&v, a->col_type);
            if (!ev_is_null(&v)) {
                ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); ev_cast_to(&r, a->col_type);
                int64_t cmp = ev_is_null(&r) ? 0 : ev_compare(&r, &v);
                if (ev_is_null(&r) || (a->agg_type == A_MIN ? cmp > 0 : cmp < 0)) memrow_set(c, dst, a->tuple_id, a->inter_slot, &v);
            }
            return;
        }
    }
}
static void agg_merge(const Ctx* c, Expr* a, const MemRow* src, MemRow* dst, int first) { /* agg_fn_call.cpp:719-822 */
    if (first) return; /* src == dst: the first row needs no merge */
    ExprValue v = memrow_get(c, src, a->tuple_id, a->inter_slot);
    if (ev_is_null(&v)) return;
    switch (a->agg_type) {
        case A_COUNT_STAR: case A_COUNT: case A_SUM: {
            ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot); ev_add(&r, &v); memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); return;
        }
        case A_AVG: {
            if (v.type != T_STRING) return;
            ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot);
            r.avg.sum += v.avg.sum; r.avg.count += v.avg.count; r.type = T_STRING;
            memrow_set(c, dst, a->tuple_id, a->inter_slot, &r); return;
        }
        case A_MIN: case A_MAX: {
            ExprValue r = memrow_get(c, dst, a->tuple_id, a->inter_slot);
            int64_t cmp = ev_is_null(&r) ? 0 : ev_compare(&r, &v);
            if (ev_is_null(&r) || (a->agg_type == A_MIN ? cmp > 0 : cmp < 0)) memrow_set(c, dst, a->tuple_id, a->inter_slot, &v);
            return;
        }
    }
}
static void agg_finalize(const Ctx* c, Expr* a, MemRow* dst) { /* agg_fn_call.cpp:927-990 */
    if (a->inter_slot == a->final_slot) return;
    if (a->agg_type == A_AVG) {
        ExprValue v = memrow_get(c, dst, a->tuple_id, a->inter_slot);
        ExprValue out = ev_null();
        if (!ev_is_null(&v) && v.avg.count != 0) { out = ev_typed(T_DOUBLE); out.u.double_val = v.avg.sum / (double)v.avg.count; }
        memrow_set(c, dst, a->tuple_id, a->final_slot, &out);
    }
}
static int agg_all_initial(const Ctx* c, Node* n, const MemRow* row) { /* AggFnCall::all_is_initialize */
    for (int i = 0; i < n->n_agg; i++) {
        Expr* a = n->aggs[i];
        if (a->is_distinct) return 0;   /* is_initialize: false for a distinct aggregate, agg_fn_call.cpp:322-328 */
        ExprValue v = memrow_get(c, row, a->tuple_id, a->inter_slot);
        if (a->agg_type == A_COUNT_STAR || a->agg_type == A_COUNT) { if (!ev_is_null(&v) && v.u.int64_val != 0) return 0; }
        else if (!ev_is_null(&v)) return 0;
    }
    return 1;
}
static MemRow agg_adopt_row(const Ctx* c, Node* n, const MemRow* src, int is_merger) {
    /* the first row of a group becomes the accumulator row (agg_node.cpp:515-531) */
    MemRow r = *src;
    const TupleDesc* t = find_tuple(c, n->agg_tuple_id);
    int ns = t ? t->n_slots : 0;
    ExprValue* slots = (ExprValue*)malloc(sizeof(ExprValue) * (size_t)(ns ? ns : 1));
    for (int i = 0; i < ns; i++) slots[i] = ev_null();
    if (is_merger && t) /* merger input rows already carry the agg tuple (as scan columns) */
        for (int i = 0; i < ns; i++) { const bko_column* col = find_col(c, n->agg_tuple_id, t->slot_ids[i]); if (col && src->idx[n->agg_tuple_id] >= 0) slots[i] = column_get(col, src->idx[n->agg_tuple_id]); }
    r.agg = slots; r.agg_tuple_id = n->agg_tuple_id;
    return r;
}
static int agg_open(Ctx* c, Node* n, int under_packet) { /* agg_node.cpp:405-505 */
    AggState* s = (AggState*)calloc(1, sizeof *s); n->st = s;
    s->is_merger = n->node_type == N_MERGE_AGG;
    Bytes key = {0, 0, 0};
    RowBatch batch; batch_init(&batch, ROW_BATCH_CAPACITY);
    int eos = 0;
    do {
        batch.n = 0;
        if (node_get_next(c, n->children[0], &batch, &eos) < 0) { batch_free(&batch); free(key.p); return -1; }
        for (int i = 0; i < batch.n; i++) { /* process_row_batch, agg_node.cpp:507-545 */
            MemRow* cur = &batch.rows[i];
            encode_exprs_key(c, n->group, n->n_group, cur, &key);
            AggEntry* e = agg_seek(s, key.p, key.n);
            int first = 0;
            if (!e) {
                MemRow adopted = agg_adopt_row(c, n, cur, s->is_merger);
                if (s->is_merger && n->n_group == 0 && agg_all_initial(c, n, &adopted)) { free(adopted.agg); continue; }
                for (int k = 0; k < n->n_agg; k++) agg_initialize(c, n->aggs[k], &adopted, 0);
                e = agg_insert(s, key.p, key.n, &adopted); first = 1;
            }
            if (s->is_merger) { for (int k = 0; k < n->n_agg; k++) { if (n->aggs[k]->is_distinct) agg_update(c, n->aggs[k], cur, &e->row); else agg_merge(c, n->aggs[k], cur, &e->row, first); } }
            else for (int k = 0; k < n->n_agg; k++) agg_update(c, n->aggs[k], cur, &e->row);
        }
    } while (!eos);
    batch_free(&batch);
    /* select count(*) from t with no rows returns 0 (agg_node.cpp:489-503) */
    if (s->n == 0 && n->n_group == 0 && (under_packet || s->is_merger)) {
        MemRow blank; memrow_init(&blank);
        MemRow adopted = agg_adopt_row(c, n, &blank, 0);
        for (int k = 0; k < n->n_agg; k++) agg_initialize(c, n->aggs[k], &adopted, 1);
        uint8_t nf = 0; agg_insert(s, &nf, 1, &adopted);
    }
    free(key.p);
    return 0;
}
static int agg_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) { /* agg_node.cpp:547-573 */
    AggState* s = (AggState*)n->st;
    for (;;) {
        if (reached_limit(n) || s->iter == s->n) { *eos = 1; return 0; }
        if (out->n >= out->cap) return 0;
        MemRow* r = &s->entries[s->iter].row;
        for (int k = 0; k < n->n_agg; k++) agg_finalize(c, n->aggs[k], r);
        out->rows[out->n++] = *r; n->num_rows_returned++; s->iter++;
    }
}

/* ---- sort ---- */
typedef struct { MemRow* rows; size_t n, cap; size_t pos; } SortState;
typedef struct { const Ctx* c; Node* n; } CmpEnv;
static CmpEnv g_cmp; /* qsort has no context argument; single-threaded like the reference bthread */
static int64_t memrow_compare(const Ctx* c, Node* n, const MemRow* l, const MemRow* r) { /* mem_row_compare.cpp:18-38 */
    for (int i = 0; i < n->n_order; i++) {
        ExprValue lv = expr_value(c, n->order[i], l), rv = expr_value(c, n->order[i], r);
        int ln = ev_is_null(&lv), rn = ev_is_null(&rv);
        if (ln && rn) continue;
        if (ln) return n->is_null_first[i] ? -1 : 1;
        if (rn) return n->is_null_first[i] ? 1 : -1;
        int64_t cmp = ev_compare(&lv, &rv);
        if (cmp != 0) return n->is_asc[i] ? cmp : -cmp;
    }
    return 0;
}
/* TopNSorter less: ties broken by arrival index (topn_sorter.h:96-106).  The full Sorter uses an
 * unstable std::sort; this restatement orders ties by arrival too, which is ONE of the orders the
 * reference may produce — tests compare full-sort results modulo tie order. */
static int sort_less_qsort(const void* a, const void* b) {
    const MemRow* l = (const MemRow*)a; const MemRow* r = (const MemRow*)b;
    int64_t cmp = memrow_compare(g_cmp.c, g_cmp.n, l, r);
    if (cmp != 0) return cmp < 0 ? -1 : 1;
    return l->arrival < r->arrival ? -1 : (l->arrival > r->arrival ? 1 : 0);
}
static int topn_less(const Ctx* c, Node* n, const MemRow* l, const MemRow* r) {
    int64_t cmp = memrow_compare(c, n, l, r);
    if (cmp != 0) return cmp < 0;
    return l->arrival < r->arrival;
}
static void heap_shiftdown(const Ctx* c, Node* n, MemRow* h, size_t cnt, size_t i) { /* topn_sorter.cpp:65-87 (max-heap) */
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < cnt && topn_less(c, n, &h[m], &h[l])) m = l;
        if (r < cnt && topn_less(c, n, &h[m], &h[r])) m = r;
        if (m == i) return;
        MemRow t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
}
static void heap_shiftup(const Ctx* c, Node* n, MemRow* h, size_t i) {
    while (i > 0) { size_t p = (i - 1) / 2; if (!topn_less(c, n, &h[p], &h[i])) return; MemRow t = h[p]; h[p] = h[i]; h[i] = t; i = p; }
}
static int sort_open(Ctx* c, Node* n) { /* sort_node.cpp:278-346 */
    SortState* s = (SortState*)calloc(1, sizeof *s); n->st = s;
    RowBatch batch; batch_init(&batch, ROW_BATCH_CAPACITY);
    int eos = 0; int64_t arrival = 0;
    do {
        batch.n = 0;
        if (node_get_next(c, n->children[0], &batch, &eos) < 0) { batch_free(&batch); return -1; }
        for (int i = 0; i < batch.n; i++) {
            MemRow row = batch.rows[i]; row.arrival = ++arrival;
            if (n->limit == -1 || (int64_t)s->n < n->limit) {
                if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 1024; s->rows = (MemRow*)realloc(s->rows, s->cap * sizeof(MemRow)); }
                s->rows[s->n++] = row;
                if (n->limit != -1) heap_shiftup(c, n, s->rows, s->n - 1);
            } else if (n->limit > 0 && topn_less(c, n, &row, &s->rows[0])) { /* topn_sorter.cpp:25-45 */
                s->rows[0] = row; heap_shiftdown(c, n, s->rows, s->n, 0);
            }
        }
    } while (!eos);
    batch_free(&batch);
    g_cmp.c = c; g_cmp.n = n;
    if (s->n > 1) qsort(s->rows, s->n, sizeof(MemRow), sort_less_qsort);
    return 0;
}
static int sort_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) {
    (void)c; SortState* s = (SortState*)n->st;
    while (out->n < out->cap && s->pos < s->n && !reached_limit(n)) { out->rows[out->n++] = s->rows[s->pos++]; n->num_rows_returned++; }
    *eos = s->pos >= s->n || reached_limit(n);
    return 0;
}

/* ---- hash join ---- */
typedef struct { uint8_t* key; size_t klen; MemRow* rows; size_t n, cap; } JoinBucket;
typedef struct {
    JoinBucket* b; size_t n, cap; int64_t* slots; size_t ns;
    Expr** outer_eq; Expr** inner_eq; int* cast_types; int n_eq;
    Expr** other; int n_other;
    unsigned char outer_tuple[MAX_TUPLES], inner_tuple[MAX_TUPLES];
    RowBatch inner; int inner_pos, inner_eos; size_t result_idx; int cur_matched;
    MemRow* outer_rows; size_t n_outer; size_t emit_outer_pos; unsigned char* outer_matched;
} JoinState;
static void collect_tuples(const Node* n, unsigned char* set) {
    if (n->node_type == N_SCAN) set[n->tuple_id] = 1;
    for (int i = 0; i < n->nchildren; i++) collect_tuples(n->children[i], set);
}
static int expr_tuple_side(const Expr* e, const unsigned char* set) { return e->node_type == E_SLOT_REF && set[e->tuple_id]; }
static void join_encode_key(const Ctx* c, Expr** eq, const int* cast, int n, const MemRow* row, Bytes* key, int* has_null) { /* joiner.cpp:608-622 */
    key->n = 0; *has_null = 0;
    for (int i = 0; i < n; i++) {
        ExprValue v = memrow_get(c, row, eq[i]->tuple_id, eq[i]->slot_id);
        if (ev_is_null(&v)) *has_null = 1;
        ev_cast_to(&v, cast[i]); key_append_value(key, &v);
    }
}
static JoinBucket* join_seek(JoinState* s, const uint8_t* key, size_t klen, int insert) {
    if ((s->n + 1) * 2 > s->ns) {
        size_t ns = s->ns ? s->ns * 2 : 1024; int64_t* sl = (int64_t*)malloc(ns * sizeof(int64_t));
        for (size_t i = 0; i < ns; i++) sl[i] = -1;
        for (size_t i = 0; i < s->n; i++) { size_t h = hash_bytes(s->b[i].key, s->b[i].klen) & (ns - 1); while (sl[h] >= 0) h = (h + 1) & (ns - 1); sl[h] = (int64_t)i; }
        free(s->slots); s->slots = sl; s->ns = ns;
    }
    size_t h = hash_bytes(key, klen) & (s->ns - 1);
    while (s->slots[h] >= 0) { JoinBucket* b = &s->b[s->slots[h]]; if (b->klen == klen && !memcmp(b->key, key, klen)) return b; h = (h + 1) & (s->ns - 1); }
    if (!insert) return NULL;
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 1024; s->b = (JoinBucket*)realloc(s->b, s->cap * sizeof(JoinBucket)); }
    JoinBucket* b = &s->b[s->n]; memset(b, 0, sizeof *b);
    b->key = (uint8_t*)malloc(klen ? klen : 1); memcpy(b->key, key, klen); b->klen = klen;
    s->slots[h] = (int64_t)s->n; s->n++;
    return b;
}
static int join_open(Ctx* c, Node* n) { /* join_node.cpp:920-1022, joiner.cpp:166-217,624-631 */
    JoinState* s = (JoinState*)calloc(1, sizeof *s); n->st = s;
    if (n->nchildren != 2) { snprintf(c->err, c->errlen, "join needs two children"); return -1; }
    if (n->join_type == J_RIGHT) { /* join_node.cpp:151-156: the right child becomes the outer (preserved, driver) table */
        Node* t = n->children[0]; n->children[0] = n->children[1]; n->children[1] = t; n->join_type = J_LEFT;
    }
    collect_tuples(n->children[0], s->outer_tuple); collect_tuples(n->children[1], s->inner_tuple);
    s->outer_eq = (Expr**)calloc((size_t)n->n_conj + 1, sizeof(Expr*)); s->inner_eq = (Expr**)calloc((size_t)n->n_conj + 1, sizeof(Expr*));
    s->cast_types = (int*)calloc((size_t)n->n_conj + 1, sizeof(int)); s->other = (Expr**)calloc((size_t)n->n_conj + 1, sizeof(Expr*));
    for (int i = 0; i < n->n_conj; i++) { /* strip_out_equal_slots */
        Expr* e = n->conj[i]; int taken = 0;
        if (e->node_type == E_FUNCTION_CALL && e->fn_op == FT_EQ && e->nchildren == 2) {
            Expr *a = e->children[0], *b = e->children[1];
            if (expr_tuple_side(a, s->outer_tuple) && expr_tuple_side(b, s->inner_tuple)) { s->outer_eq[s->n_eq] = a; s->inner_eq[s->n_eq] = b; taken = 1; }
            else if (expr_tuple_side(b, s->outer_tuple) && expr_tuple_side(a, s->inner_tuple)) { s->outer_eq[s->n_eq] = b; s->inner_eq[s->n_eq] = a; taken = 1; }
            if (taken) {
                int ot = s->outer_eq[s->n_eq]->col_type, it = s->inner_eq[s->n_eq]->col_type;
                if (ot == it) s->cast_types[s->n_eq] = ot;
                else if (is_signed(ot) && is_signed(it)) s->cast_types[s->n_eq] = T_INT64;
                else if (is_uint(ot) && is_uint(it)) s->cast_types[s->n_eq] = T_UINT64;
                else { snprintf(c->err, c->errlen, "join key cast to STRING is outside the path"); return -1; }
                s->n_eq++;
            }
        }
        if (!taken) s->other[s->n_other++] = e;
    }
    if (s->n_eq == 0) { snprintf(c->err, c->errlen, "join without an equality condition"); return -1; }
    /* fetch the whole outer (driver) table and build the map */
    RowBatch batch; batch_init(&batch, ROW_BATCH_CAPACITY); int eos = 0; Bytes key = {0, 0, 0};
    size_t ocap = 0;
    do {
        batch.n = 0;
        if (node_get_next(c, n->children[0], &batch, &eos) < 0) { batch_free(&batch); return -1; }
        for (int i = 0; i < batch.n; i++) {
            if (s->n_outer == ocap) { ocap = ocap ? ocap * 2 : 1024; s->outer_rows = (MemRow*)realloc(s->outer_rows, ocap * sizeof(MemRow)); }
            s->outer_rows[s->n_outer] = batch.rows[i]; s->outer_rows[s->n_outer].arrival = (int64_t)s->n_outer; s->n_outer++;
        }
    } while (!eos);
    for (size_t i = 0; i < s->n_outer; i++) {
        int has_null; join_encode_key(c, s->outer_eq, s->cast_types, s->n_eq, &s->outer_rows[i], &key, &has_null);
        /* SQL semantics: a NULL key never matches.  (The row engine encodes NULL as empty bytes,
         * mut_table_key.h:167-199, so NULL==NULL there; Acero's hashjoin and MySQL do not match
         * NULLs.  SURVEY.md Appendix B item 13: the GPU path follows SQL/Acero.) */
        if (has_null) continue;
        JoinBucket* b = join_seek(s, key.p, key.n, 1);
        if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 2; b->rows = (MemRow*)realloc(b->rows, b->cap * sizeof(MemRow)); }
        b->rows[b->n++] = s->outer_rows[i];
    }
    s->outer_matched = (unsigned char*)calloc(s->n_outer + 1, 1);
    free(key.p); batch_free(&batch);
    batch_init(&s->inner, ROW_BATCH_CAPACITY);
    return 0;
}
static int join_satisfy(const Ctx* c, JoinState* s, const MemRow* row) { /* joiner.cpp:598-606 */
    for (int i = 0; i < s->n_other; i++) { ExprValue v = expr_value(c, s->other[i], row); if (ev_is_null(&v) || !num_bool(&v)) return 0; }
    return 1;
}
static int join_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) { /* join_node.cpp:1277-1326 (+ LEFT/SEMI/ANTI on the outer side) */
    JoinState* s = (JoinState*)n->st; Bytes key = {0, 0, 0};
    for (;;) {
        if (s->inner_pos >= s->inner.n) {
            if (s->inner_eos) break;
            s->inner.n = 0; s->inner_pos = 0;
            if (node_get_next(c, n->children[1], &s->inner, &s->inner_eos) < 0) { free(key.p); return -1; }
            continue;
        }
        MemRow* in = &s->inner.rows[s->inner_pos];
        int has_null; join_encode_key(c, s->inner_eq, s->cast_types, s->n_eq, in, &key, &has_null);
        JoinBucket* b = has_null ? NULL : join_seek(s, key.p, key.n, 0);
        if (b) {
            for (; s->result_idx < b->n; s->result_idx++) {
                if (reached_limit(n)) { *eos = 1; free(key.p); return 0; }
                if (out->n >= out->cap) { free(key.p); return 0; }
                MemRow merged = b->rows[s->result_idx]; /* construct_result_batch: copy both sides */
                for (int t = 0; t < MAX_TUPLES; t++) if (s->inner_tuple[t]) merged.idx[t] = in->idx[t];
                if (join_satisfy(c, s, &merged)) {
                    s->outer_matched[b->rows[s->result_idx].arrival] = 1;
                    if (n->join_type == J_INNER || n->join_type == J_LEFT) { out->rows[out->n++] = merged; n->num_rows_returned++; }
                }
            }
        }
        s->result_idx = 0; s->inner_pos++;
    }
    /* inner exhausted: LEFT emits unmatched outer rows NULL-extended; SEMI/ANTI emit outer rows */
    if (n->join_type == J_LEFT || n->join_type == J_SEMI || n->join_type == J_ANTI) {
        for (; s->emit_outer_pos < s->n_outer; s->emit_outer_pos++) {
            if (reached_limit(n)) break;
            if (out->n >= out->cap) { free(key.p); return 0; }
            int m = s->outer_matched[s->emit_outer_pos];
            if ((n->join_type == J_LEFT && !m) || (n->join_type == J_SEMI && m) || (n->join_type == J_ANTI && !m)) { out->rows[out->n++] = s->outer_rows[s->emit_outer_pos]; n->num_rows_returned++; }
        }
    }
    *eos = 1; free(key.p); return 0;
}

/* ---- limit ---- */
typedef struct { int64_t skipped; RowBatch child; int child_pos, child_eos; } LimitState;
static int limit_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) { /* limit_node.cpp:21-134: skip `offset`, pass `limit` */
    LimitState* s = (LimitState*)n->st;
    for (;;) {
        if (reached_limit(n)) { *eos = 1; return 0; }
        if (out->n >= out->cap) return 0;
        if (s->child_pos >= s->child.n) {
            if (s->child_eos) { *eos = 1; return 0; }
            s->child.n = 0; s->child_pos = 0;
            if (node_get_next(c, n->children[0], &s->child, &s->child_eos) < 0) return -1;
            continue;
        }
        MemRow* row = &s->child.rows[s->child_pos++];
        if (s->skipped < n->offset) { s->skipped++; continue; }
        out->rows[out->n++] = *row; n->num_rows_returned++;
    }
}

static int has_ancestor_packet = 0;
static int node_open(Ctx* c, Node* n) {
    int was_packet = has_ancestor_packet;
    if (n->node_type == N_PACKET) has_ancestor_packet = 1;
    int r = 0;
    /* ExecNode::open opens children first (exec_node.cpp:315-330); AGG/SORT/JOIN then drain them */
    for (int i = 0; i < n->nchildren && r == 0; i++) r = node_open(c, n->children[i]);
    if (r == 0) switch (n->node_type) {
        case N_SCAN: r = scan_open(c, n); break;
        case N_WHERE_FILTER: case N_TABLE_FILTER: case N_HAVING_FILTER: {
            FilterState* s = (FilterState*)calloc(1, sizeof *s); batch_init(&s->child, ROW_BATCH_CAPACITY); n->st = s; break; }
        case N_AGG: case N_MERGE_AGG: r = agg_open(c, n, has_ancestor_packet); break;
        case N_SORT: r = sort_open(c, n); break;
        case N_JOIN: r = join_open(c, n); break;
        case N_LIMIT: { LimitState* s = (LimitState*)calloc(1, sizeof *s); batch_init(&s->child, ROW_BATCH_CAPACITY); n->st = s; break; }
        default: break;
    }
    has_ancestor_packet = was_packet;
    return r;
}
static int node_get_next(Ctx* c, Node* n, RowBatch* out, int* eos) {
    switch (n->node_type) {
        case N_SCAN: return scan_get_next(c, n, out, eos);
        case N_WHERE_FILTER: case N_TABLE_FILTER: case N_HAVING_FILTER: return filter_get_next(c, n, out, eos);
        case N_AGG: case N_MERGE_AGG: return agg_get_next(c, n, out, eos);
        case N_SORT: return sort_get_next(c, n, out, eos);
        case N_JOIN: return join_get_next(c, n, out, eos);
        case N_LIMIT: return limit_get_next(c, n, out, eos);
        case N_PACKET: case N_SELECT_MANAGER: return node_get_next(c, n->children[0], out, eos);
        default: return -1;
    }
}

static int infer_node(Ctx* c, Node* n) {
    for (int i = 0; i < n->n_conj; i++) if (type_infer(c, n->conj[i]) < 0) return -1;
    for (int i = 0; i < n->n_group; i++) if (type_infer(c, n->group[i]) < 0) return -1;
    for (int i = 0; i < n->n_agg; i++) if (type_infer(c, n->aggs[i]) < 0) return -1;
    for (int i = 0; i < n->n_order; i++) if (type_infer(c, n->order[i]) < 0) return -1;
    for (int i = 0; i < n->nchildren; i++) if (infer_node(c, n->children[i]) < 0) return -1;
    return 0;
}

/* =========================== result materialisation =========================== */
typedef struct { int tuple_id, slot_id, type, is_blob; Expr* expr; } OutCol;
static int elem_bytes(int storage) { switch (storage) { case T_INT32: case T_UINT32: case T_FLOAT: return 4; case T_BOOL: return 1; case T_STRING: return 16; default: return 8; } }

static Node* find_top_shaper(Node* n) { /* the node that determines the output tuple set */
    while (n && (n->node_type == N_PACKET || n->node_type == N_SELECT_MANAGER || n->node_type == N_LIMIT ||
                 n->node_type == N_SORT || n->node_type == N_HAVING_FILTER)) n = n->nchildren ? n->children[0] : NULL;
    return n;
}
static void collect_scan_tuples(const Node* n, int* ids, int* cnt) {
    if (n->node_type == N_SCAN) ids[(*cnt)++] = n->tuple_id;
    for (int i = 0; i < n->nchildren; i++) collect_scan_tuples(n->children[i], ids, cnt);
}

int bko_execute(const uint8_t* plan, size_t len, const bko_column* in_cols, int n_in, bko_result** out, char* err, size_t errlen) {
    Ctx c; memset(&c, 0, sizeof c); c.cols = in_cols; c.n_cols = n_in; c.err = err; c.errlen = errlen;
    if (err && errlen) err[0] = 0;
    Reader r; memset(&r, 0, sizeof r); r.w = (const int32_t*)plan; r.n = len / 4; r.err = err; r.errlen = errlen;
    if ((uint32_t)rd(&r) != 0x31504B42u || rd(&r) != 1) { set_err(&r, "bad plan magic/version"); return -1; }
    c.n_tuples = rd(&r); int n_nodes = rd(&r);
    if (c.n_tuples < 0 || c.n_tuples > MAX_TUPLES) { set_err(&r, "too many tuples"); return -1; }
    for (int i = 0; i < c.n_tuples; i++) {
        TupleDesc* t = &c.tuples[i]; t->tuple_id = rd(&r); t->n_slots = rd(&r);
        if (t->tuple_id < 0 || t->tuple_id >= MAX_TUPLES || t->n_slots < 0 || t->n_slots > 4096) { set_err(&r, "bad tuple descriptor"); return -1; }
        t->slot_ids = (int*)calloc((size_t)t->n_slots + 1, sizeof(int)); t->types = (int*)calloc((size_t)t->n_slots + 1, sizeof(int));
        for (int k = 0; k < t->n_slots; k++) { t->slot_ids[k] = rd(&r); t->types[k] = rd(&r); }
    }
    Node* root = parse_node(&r, &n_nodes);
    if (r.fail || !root) { if (err && !err[0]) snprintf(err, errlen, "truncated plan"); return -1; }
    if (infer_node(&c, root) < 0) return -2;
    has_ancestor_packet = 0;
    { int ids[MAX_TUPLES], cnt = 0; collect_scan_tuples(root, ids, &cnt); g_scan_nodes = cnt; }
    if (node_open(&c, root) < 0) return -3;

    /* output schema */
    Node* shaper = find_top_shaper(root);
    OutCol oc[256]; int noc = 0;
    if (shaper && (shaper->node_type == N_AGG || shaper->node_type == N_MERGE_AGG)) {
        for (int i = 0; i < shaper->n_group; i++) {
            Expr* g = shaper->group[i];
            OutCol o = { g->node_type == E_SLOT_REF ? g->tuple_id : -1, g->node_type == E_SLOT_REF ? g->slot_id : i, g->col_type, 0, g };
            oc[noc++] = o;
        }
        for (int i = 0; i < shaper->n_agg; i++) {
            Expr* a = shaper->aggs[i];
            if (a->inter_slot != a->final_slot) { OutCol o = { a->tuple_id, a->inter_slot, T_STRING, 1, NULL }; oc[noc++] = o; }
            OutCol o = { a->tuple_id, a->final_slot, slot_type(&c, a->tuple_id, a->final_slot), 0, NULL };
            if (o.type == T_INVALID) o.type = a->col_type;
            oc[noc++] = o;
        }
    } else {
        int ids[MAX_TUPLES], cnt = 0; collect_scan_tuples(root, ids, &cnt);
        for (int t = 0; t < cnt; t++) for (int i = 0; i < n_in; i++) if (in_cols[i].tuple_id == ids[t]) {
            OutCol o = { in_cols[i].tuple_id, in_cols[i].slot_id, in_cols[i].prim_type, in_cols[i].prim_type == T_STRING, NULL };
            oc[noc++] = o;
        }
    }
    /* drain */
    size_t cap = 1024, nrows = 0; MemRow* rows = (MemRow*)malloc(cap * sizeof(MemRow));
    RowBatch batch; batch_init(&batch, ROW_BATCH_CAPACITY); int eos = 0;
    do {
        batch.n = 0;
        if (node_get_next(&c, root, &batch, &eos) < 0) return -4;
        for (int i = 0; i < batch.n; i++) { if (nrows == cap) { cap *= 2; rows = (MemRow*)realloc(rows, cap * sizeof(MemRow)); } rows[nrows++] = batch.rows[i]; }
    } while (!eos);
    batch_free(&batch);
    bko_result* res = (bko_result*)calloc(1, sizeof *res);
    res->ncols = noc; res->nrows = (int64_t)nrows; res->cols = (bko_column*)calloc((size_t)noc + 1, sizeof(bko_column));
    res->rows_scanned = c.rows_scanned; res->rows_filtered = c.rows_filtered;
    for (int k = 0; k < noc; k++) {
        int st = oc[k].is_blob ? T_STRING : storage_type(oc[k].type); int eb = elem_bytes(st);
        uint8_t* vals = (uint8_t*)calloc(nrows ? nrows : 1, (size_t)eb);
        uint8_t* valid = (uint8_t*)calloc((nrows + 7) / 8 + 1, 1);
        for (size_t i = 0; i < nrows; i++) {
            ExprValue v;
            if (oc[k].expr && oc[k].tuple_id < 0) v = expr_value(&c, oc[k].expr, &rows[i]);
            else v = memrow_get(&c, &rows[i], oc[k].tuple_id, oc[k].slot_id);
            if (ev_is_null(&v)) continue;
            valid[i >> 3] |= (uint8_t)(1u << (i & 7));
            if (st == T_STRING) { memcpy(vals + 16 * i, &v.avg, 16); continue; }
            ev_cast_to(&v, st);
            memcpy(vals + (size_t)eb * i, &v.u, (size_t)eb);
        }
        bko_column* oc_out = &res->cols[k];
        oc_out->tuple_id = oc[k].tuple_id; oc_out->slot_id = oc[k].slot_id; oc_out->prim_type = oc[k].is_blob ? T_STRING : oc[k].type;
        oc_out->elem_size = eb; oc_out->values = vals; oc_out->validity = valid; oc_out->length = (int64_t)nrows;
    }
    free(rows);
    *out = res;
    return 0; /* plan/state memory is intentionally leaked to the process arena: the oracle is a
                 short-lived checker, never a service */
}

void bko_free_result(bko_result* r) {
    if (!r) return;
    for (int i = 0; i < r->ncols; i++) { free((void*)r->cols[i].values); free((void*)r->cols[i].validity); }
    free(r->cols); free(r);
}
