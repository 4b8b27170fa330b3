// oracle/flopcount.h -- TEST TOOLING.  Counting double: every arithmetic operation of the oracle increments a counter.
#include <cmath>
#include <cstring>
#include <cstdlib>
extern "C" { extern long long o_flops[8]; }   // 0 add/sub, 1 mul, 2 div, 3 sqrt, 4 transcendental (sin cos pow), 5 fabs/min/max/neg (not counted as FLOP)
struct CD {
  double v;
  CD() = default;
  CD(double x) : v(x) {}
  CD(int x) : v(x) {}
  CD(long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  explicit operator bool() const { return v != 0; }
  CD& operator+=(CD o) { o_flops[0]++; v += o.v; return *this; }
  CD& operator-=(CD o) { o_flops[0]++; v -= o.v; return *this; }
  CD& operator*=(CD o) { o_flops[1]++; v *= o.v; return *this; }
  CD& operator/=(CD o) { o_flops[2]++; v /= o.v; return *this; }
};
static inline CD operator+(CD a, CD b) { o_flops[0]++; return CD(a.v + b.v); }
static inline CD operator-(CD a, CD b) { o_flops[0]++; return CD(a.v - b.v); }
static inline CD operator*(CD a, CD b) { o_flops[1]++; return CD(a.v * b.v); }
static inline CD operator/(CD a, CD b) { o_flops[2]++; return CD(a.v / b.v); }
static inline CD operator-(CD a) { return CD(-a.v); }
static inline CD operator+(CD a) { return a; }
#define MIX(op) \
  static inline CD operator op(CD a, double b) { return a op CD(b); } static inline CD operator op(double a, CD b) { return CD(a) op b; } \
  static inline CD operator op(CD a, int b) { return a op CD(b); } static inline CD operator op(int a, CD b) { return CD(a) op b; }
MIX(+) MIX(-) MIX(*) MIX(/)
#define CMP(op) static inline bool operator op(CD a, CD b) { return a.v op b.v; } static inline bool operator op(CD a, double b) { return a.v op b; } \
  static inline bool operator op(double a, CD b) { return a op b.v; } static inline bool operator op(CD a, int b) { return a.v op b; } static inline bool operator op(int a, CD b) { return a op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
static inline CD sqrt(CD a) { o_flops[3]++; return CD(std::sqrt(a.v)); }
static inline CD sin(CD a) { o_flops[4]++; return CD(std::sin(a.v)); }
static inline CD cos(CD a) { o_flops[4]++; return CD(std::cos(a.v)); }
static inline CD pow(CD a, CD b) { o_flops[4]++; return CD(std::pow(a.v, b.v)); }
static inline CD pow(CD a, double b) { o_flops[4]++; return CD(std::pow(a.v, b)); }
static inline CD pow(double a, CD b) { o_flops[4]++; return CD(std::pow(a, b.v)); }
static inline CD fabs(CD a) { o_flops[5]++; return CD(std::fabs(a.v)); }
static inline CD fmin(CD a, CD b) { o_flops[5]++; return CD(std::fmin(a.v, b.v)); }
static inline CD fmax(CD a, CD b) { o_flops[5]++; return CD(std::fmax(a.v, b.v)); }
static inline CD fmin(CD a, double b) { return fmin(a, CD(b)); } static inline CD fmin(double a, CD b) { return fmin(CD(a), b); }
static inline CD fmax(CD a, double b) { return fmax(a, CD(b)); } static inline CD fmax(double a, CD b) { return fmax(CD(a), b); }
static inline CD fmax(int a, CD b) { return fmax(CD(a), b); } static inline CD fmin(int a, CD b) { return fmin(CD(a), b); }
static inline CD fmax(CD a, int b) { return fmax(a, CD(b)); } static inline CD fmin(CD a, int b) { return fmin(a, CD(b)); }
#define double CD
