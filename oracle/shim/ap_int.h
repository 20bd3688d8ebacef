// ap_int.h - minimal stand-in for Xilinx Vivado-HLS "ap_int.h", which the reference's
// hls/ sources include (hls/windows/win_function.h:45, hls/cordic/cordic.h:44) but which is a
// third-party header not present under /root/reference (no version pinned anywhere in the repo).
//
// TEST INFRASTRUCTURE ONLY: used to compile the UNMODIFIED reference sources into oracle/_ref/.
// It restates the published arbitrary-precision semantics for exactly the operations those two
// files use:
//   * ap_int<W>/ap_uint<W> hold a W-bit two's-complement / unsigned value; construction and
//     assignment from anything wider wrap modulo 2^W (doubles truncate toward zero first);
//   * + - * & | ^ between ap values and/or C integers are exact (result wide enough);
//   * ap_int<W> >> n and << n keep width W (>> arithmetic for signed, << drops the top bits);
//   * ~x keeps width W; comparisons compare values.
// Exact intermediates are carried in __int128, enough for every expression in the reference
// (the widest is a (2*NWIDTH+1)-bit coefficient times an NWIDTH-bit cosine, NWIDTH <= 32).
#ifndef BHW_ORACLE_AP_INT_SHIM_H_
#define BHW_ORACLE_AP_INT_SHIM_H_
#include <type_traits>

struct ap_wide {
  __int128 v;
  ap_wide() : v(0) {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_wide(T x) : v((__int128)x) {}
  static ap_wide raw(__int128 x) { ap_wide r; r.v = x; return r; }
  ap_wide operator>>(int n) const { return raw(v >> n); }
  ap_wide operator<<(int n) const { return raw((__int128)((unsigned __int128)v << n)); }
  ap_wide operator~() const { return raw(~v); }
  ap_wide operator-() const { return raw(-v); }
  long long to_int64() const { return (long long)v; }
};
inline ap_wide operator+(ap_wide a, ap_wide b) { return ap_wide::raw(a.v + b.v); }
inline ap_wide operator-(ap_wide a, ap_wide b) { return ap_wide::raw(a.v - b.v); }
inline ap_wide operator*(ap_wide a, ap_wide b) { return ap_wide::raw(a.v * b.v); }
inline ap_wide operator&(ap_wide a, ap_wide b) { return ap_wide::raw(a.v & b.v); }
inline ap_wide operator|(ap_wide a, ap_wide b) { return ap_wide::raw(a.v | b.v); }
inline ap_wide operator^(ap_wide a, ap_wide b) { return ap_wide::raw(a.v ^ b.v); }
inline bool operator<(ap_wide a, ap_wide b) { return a.v < b.v; }
inline bool operator>(ap_wide a, ap_wide b) { return a.v > b.v; }
inline bool operator<=(ap_wide a, ap_wide b) { return a.v <= b.v; }
inline bool operator>=(ap_wide a, ap_wide b) { return a.v >= b.v; }
inline bool operator==(ap_wide a, ap_wide b) { return a.v == b.v; }
inline bool operator!=(ap_wide a, ap_wide b) { return a.v != b.v; }

template <int W, bool S>
struct ap_base {
  static_assert(W >= 1 && W <= 120, "width out of range for this stand-in");
  __int128 v;  // canonical: sign-extended (S) or zero-extended (!S) W-bit value
  static __int128 wrap(__int128 x) {
    const unsigned __int128 mask = (((unsigned __int128)1) << W) - 1;
    unsigned __int128 u = (unsigned __int128)x & mask;
    if (S) {
      const unsigned __int128 sign = ((unsigned __int128)1) << (W - 1);
      return (__int128)((u ^ sign) - sign);
    }
    return (__int128)u;
  }
  ap_base() : v(0) {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_base(T x) : v(wrap((__int128)x)) {}
  ap_base(ap_wide x) : v(wrap(x.v)) {}
  template <int W2, bool S2>
  ap_base(const ap_base<W2, S2>& o) : v(wrap(o.v)) {}
  operator ap_wide() const { return ap_wide::raw(v); }
  ap_base operator>>(int n) const { ap_base r; r.v = wrap(v >> n); return r; }
  ap_base operator<<(int n) const { ap_base r; r.v = wrap((__int128)((unsigned __int128)v << n)); return r; }
  ap_base operator~() const { ap_base r; r.v = wrap(~v); return r; }
  long long to_int64() const { return (long long)v; }
  explicit operator long long() const { return (long long)v; }
  explicit operator int() const { return (int)v; }
  explicit operator double() const { return (double)v; }
};

template <int W> struct ap_int : ap_base<W, true> {
  typedef ap_base<W, true> B;
  ap_int() {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_int(T x) : B(x) {}
  ap_int(ap_wide x) : B(x) {}
  template <int W2, bool S2> ap_int(const ap_base<W2, S2>& o) : B(o) {}
};
template <int W> struct ap_uint : ap_base<W, false> {
  typedef ap_base<W, false> B;
  ap_uint() {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  ap_uint(T x) : B(x) {}
  ap_uint(ap_wide x) : B(x) {}
  template <int W2, bool S2> ap_uint(const ap_base<W2, S2>& o) : B(o) {}
};
#endif
