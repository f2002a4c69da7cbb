// orc_flopcount.hpp — counting scalar for the ORC_COUNT_FLOPS build of the oracle.  TEST INFRASTRUCTURE ONLY.
//
// rbd_oracle.c compiled as C++ with -DORC_COUNT_FLOPS sees `double` as this type (a struct holding one double, same
// size and layout, so the C entry points keep their binary interface): every addition / subtraction, multiplication,
// division, square root and sine / cosine evaluation the restatement performs is counted per thread.  The counts are
// the FP64 work of the reference algorithm as restated (pinocchio's generic code path: dense joint subspaces, SE3
// actions per joint) — SURVEY.md §8(d) "algorithmic flops / eval".
#pragma once
#include <math.h>

struct orc_counters { long long add, mul, div, sqrt_, sincos, cmp; };
extern thread_local orc_counters orc_cnt;

struct orc_counted {
  double v;
  orc_counted() = default;
  orc_counted(double x) : v(x) {}
  orc_counted(int x) : v(x) {}
  orc_counted(long x) : v((double)x) {}
  orc_counted(long double x) : v((double)x) {}  // <float.h> spells DBL_EPSILON as double(...L)
};
inline orc_counted operator+(orc_counted a, orc_counted b) { ++orc_cnt.add; return orc_counted(a.v + b.v); }
inline orc_counted operator-(orc_counted a, orc_counted b) { ++orc_cnt.add; return orc_counted(a.v - b.v); }
inline orc_counted operator*(orc_counted a, orc_counted b) { ++orc_cnt.mul; return orc_counted(a.v * b.v); }
inline orc_counted operator/(orc_counted a, orc_counted b) { ++orc_cnt.div; return orc_counted(a.v / b.v); }
inline orc_counted operator-(orc_counted a) { return orc_counted(-a.v); }
inline orc_counted& operator+=(orc_counted& a, orc_counted b) { ++orc_cnt.add; a.v += b.v; return a; }
inline orc_counted& operator-=(orc_counted& a, orc_counted b) { ++orc_cnt.add; a.v -= b.v; return a; }
inline orc_counted& operator*=(orc_counted& a, orc_counted b) { ++orc_cnt.mul; a.v *= b.v; return a; }
inline bool operator>(orc_counted a, orc_counted b) { ++orc_cnt.cmp; return a.v > b.v; }
inline bool operator<(orc_counted a, orc_counted b) { ++orc_cnt.cmp; return a.v < b.v; }
inline bool operator>=(orc_counted a, orc_counted b) { ++orc_cnt.cmp; return a.v >= b.v; }
inline bool operator<=(orc_counted a, orc_counted b) { ++orc_cnt.cmp; return a.v <= b.v; }
inline orc_counted fabs(orc_counted a) { return orc_counted(::fabs(a.v)); }
inline orc_counted sqrt(orc_counted a) { ++orc_cnt.sqrt_; return orc_counted(::sqrt(a.v)); }
inline orc_counted cos(orc_counted a) { ++orc_cnt.sincos; return orc_counted(::cos(a.v)); }
inline orc_counted sin(orc_counted a) { ++orc_cnt.sincos; return orc_counted(::sin(a.v)); }
