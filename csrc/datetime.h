// datetime.h — the reference's DATE / DATETIME / TIMESTAMP / TIME images and the conversions between them, usable from host
// code (plan.cpp folds literals) and from kernels (OP_CAST on a column).
//
// Layouts (reference include/common/datetime.h:28-33,56-68; src/common/datetime.cpp:248-262):
//   DATETIME  u64  (year*13 + month) << 46 | day << 41 | hour << 36 | minute << 30 | second << 24 | microsecond
//   DATE      u32  DATETIME >> 41
//   TIMESTAMP u32  seconds since 1970-01-01 00:00:00 UTC; the reference converts in a fixed UTC+8 zone without DST
//                  (mktime_fixed_r / localtime_fixed_r, datetime.cpp:41-98, default tz_offset_hours = 8)
//   TIME      i32  sign * (hour << 12 | minute << 6 | second)
// All four order the way their images order (unsigned for the first three, signed for TIME), so comparisons and MIN/MAX work
// on the raw image; only a change of type needs the functions below (ExprValue::cast_to, include/common/expr_value.h:534-573).
// The calendar arithmetic is Richards' Julian-day-number form, checked against the oracle's restatement of the reference over
// every day of years 0..9999 and over the whole uint32 TIMESTAMP range (tests/test_datetime.py).
#pragma once
#include <stdint.h>
#include "../include/bkgpu_plan.h"

#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#define BK_HD inline
#endif

namespace bk {

constexpr int64_t DT_ZONE_SECONDS = 8 * 3600;      // the reference's fixed zone
constexpr int64_t DT_EPOCH_JDN = 2440588;          // Julian day number of 1970-01-01

BK_HD uint64_t dt_make(uint64_t year, uint64_t month, uint64_t day, uint64_t hour, uint64_t minute, uint64_t second, uint64_t micro) {
    return ((year * 13 + month) << 46) | (day << 41) | (hour << 36) | (minute << 30) | (second << 24) | micro;
}
// days since the epoch of (year >= 0, month 1..12, day); `day` may run past the month's end, it adds linearly (as in the reference)
BK_HD int64_t dt_epoch_days(int64_t year, int64_t month, int64_t day) {
    const int64_t a = (14 - month) / 12, y = year + 4800 - a, m = month + 12 * a - 3;
    return day + (153 * m + 2) / 5 + 365 * y + y / 4 - y / 100 + y / 400 - 32045 - DT_EPOCH_JDN;
}
// DATETIME -> TIMESTAMP (datetime_to_timestamp, datetime.cpp:304-331): 0 for the zero date, a zero month or day, and anything at
// or before the epoch; the caller stores the low 32 bits (expr_value.h:551)
BK_HD int64_t dt_datetime_to_timestamp(uint64_t dt) {
    if (dt == 0) return 0;
    const int64_t ym = (int64_t)((dt >> 46) & 0x1FFFF), year = ym / 13, month = ym % 13, day = (int64_t)((dt >> 41) & 0x1F);
    if (month == 0 || day == 0) return 0;
    const int64_t t = dt_epoch_days(year, month, day) * 86400 + (int64_t)((dt >> 36) & 0x1F) * 3600 + (int64_t)((dt >> 30) & 0x3F) * 60 +
                      (int64_t)((dt >> 24) & 0x3F) - DT_ZONE_SECONDS;
    return t <= 0 ? 0 : t;
}
// TIMESTAMP -> DATETIME (timestamp_to_datetime, datetime.cpp:352-373); ts > 0 here (uint32 source), microseconds are zero
BK_HD uint64_t dt_timestamp_to_datetime(int64_t ts) {
    if (ts == 0) return 0;
    const int64_t t = ts + DT_ZONE_SECONDS;
    int64_t days = t / 86400, rem = t % 86400;
    if (rem < 0) { rem += 86400; days--; }
    const int64_t J = days + DT_EPOCH_JDN;
    const int64_t f = J + 1401 + (((4 * J + 274277) / 146097) * 3) / 4 - 38, e = 4 * f + 3, g = (e % 1461) / 4, h = 5 * g + 2;
    const int64_t day = (h % 153) / 5 + 1, month = ((h / 153 + 2) % 12) + 1, year = e / 1461 - 4716 + (14 - month) / 12;
    return dt_make((uint64_t)year, (uint64_t)month, (uint64_t)day, (uint64_t)(rem / 3600), (uint64_t)((rem % 3600) / 60), (uint64_t)(rem % 60), 0);
}
BK_HD uint64_t dt_date_to_datetime(uint64_t date) { return date << 41; }                 // datetime.h:65-67
BK_HD uint64_t dt_datetime_to_date(uint64_t dt) { return (dt >> 41) & 0x3FFFFF; }         // datetime.h:62-64
BK_HD int64_t dt_datetime_to_time(uint64_t dt) {                                          // datetime_to_time, datetime.cpp:410-419
    return (int64_t)(((dt >> 24) & 0x3F) | (((dt >> 30) & 0x3F) << 6) | (((dt >> 36) & 0x1F) << 12));
}
BK_HD bool dt_is_family(int prim) { return prim == BK_DATETIME || prim == BK_TIMESTAMP || prim == BK_DATE || prim == BK_TIME; }
// A change of type inside the family (canonical 64-bit images in and out: TIMESTAMP / DATE zero-extended, TIME sign-extended).
// TIME as the source is relative to the current date in the reference (time_to_datetime, datetime.cpp:420-442): plans that need it
// are rejected when they are compiled, so `from` is never BK_TIME here.
BK_HD uint64_t dt_family_cast(uint64_t v, int from, int to) {
    if (from == to) return v;
    const uint64_t dt = from == BK_DATETIME ? v : (from == BK_TIMESTAMP ? dt_timestamp_to_datetime((int64_t)(uint32_t)v) : dt_date_to_datetime((uint32_t)v));
    switch (to) {
        case BK_TIMESTAMP: return (uint64_t)(uint32_t)dt_datetime_to_timestamp(dt);
        case BK_DATE: return dt_datetime_to_date(dt);
        case BK_TIME: return (uint64_t)dt_datetime_to_time(dt);
        default: return dt;
    }
}

}  // namespace bk
