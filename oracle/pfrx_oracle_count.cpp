// pfrx_oracle_count.cpp -- OP-COUNTING build of the oracle (test / measurement infrastructure, like
// pfrx_oracle.c itself: only tests/ and bench.py's roofline bookkeeping load it).
//
// SURVEY.md section 8(d) defines the algorithmic flops of a cell-solve as what the REFERENCE algorithm
// executes, counted with add / sub / mul / div / compare-select = 1 (an FMA is a multiply and an add:
// 2) and each exp / log / pow / sqrt / atan = 20.  This translation unit compiles the very same source,
// pfrx_oracle.c, as C++ with `double` replaced by a scalar whose operators count, so the number is
// the oracle's own operation stream on the actual cells -- not a closed form.  The C ABI structs keep
// real doubles (pfrx.h is included first); conversions to and from them are not counted.
//
//   g++ -O1 -shared -fPIC -fpermissive -w -o _build/libpfrx_oracle_count.so pfrx_oracle_count.cpp -lm -lpthread
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pfrx.h"

static __thread unsigned long long g_ops = 0;
static unsigned long long g_ops_total = 0;
static pthread_mutex_t g_ops_mu = PTHREAD_MUTEX_INITIALIZER;

struct Cd {
  double v;
  Cd() = default;
  Cd(double x) : v(x) {}
  Cd(int x) : v(x) {}
  Cd(long x) : v((double)x) {}
  Cd(long long x) : v((double)x) {}
  Cd(unsigned x) : v(x) {}
  Cd(unsigned long x) : v((double)x) {}
  Cd(float x) : v(x) {}
  operator double() const { return v; }
  Cd &operator+=(Cd o) { g_ops++; v += o.v; return *this; }
  Cd &operator-=(Cd o) { g_ops++; v -= o.v; return *this; }
  Cd &operator*=(Cd o) { g_ops++; v *= o.v; return *this; }
  Cd &operator/=(Cd o) { g_ops++; v /= o.v; return *this; }
  Cd operator-() const { return Cd(-v); }
  Cd operator+() const { return *this; }
};
#define CD_BIN(op)                                                                          \
  static inline Cd operator op(Cd a, Cd b) { g_ops++; return Cd(a.v op b.v); }              \
  static inline Cd operator op(Cd a, double b) { g_ops++; return Cd(a.v op b); }            \
  static inline Cd operator op(double a, Cd b) { g_ops++; return Cd(a op b.v); }            \
  static inline Cd operator op(Cd a, int b) { g_ops++; return Cd(a.v op b); }               \
  static inline Cd operator op(int a, Cd b) { g_ops++; return Cd(a op b.v); }
CD_BIN(+) CD_BIN(-) CD_BIN(*) CD_BIN(/)
#define CD_CMP(op)                                                                          \
  static inline bool operator op(Cd a, Cd b) { g_ops++; return a.v op b.v; }                \
  static inline bool operator op(Cd a, double b) { g_ops++; return a.v op b; }              \
  static inline bool operator op(double a, Cd b) { g_ops++; return a op b.v; }              \
  static inline bool operator op(Cd a, int b) { g_ops++; return a.v op b; }                 \
  static inline bool operator op(int a, Cd b) { g_ops++; return a op b.v; }
CD_CMP(<) CD_CMP(>) CD_CMP(<=) CD_CMP(>=) CD_CMP(==) CD_CMP(!=)
#define CD_TR1(f) static inline Cd f(Cd a) { g_ops += 20; return Cd(::f(a.v)); }
CD_TR1(exp) CD_TR1(log) CD_TR1(log10) CD_TR1(sqrt) CD_TR1(atan)
static inline Cd pow(Cd a, Cd b) { g_ops += 20; return Cd(::pow(a.v, b.v)); }
static inline Cd pow(Cd a, double b) { g_ops += 20; return Cd(::pow(a.v, b)); }
static inline Cd pow(double a, Cd b) { g_ops += 20; return Cd(::pow(a, b.v)); }
static inline Cd pow(Cd a, int b) { g_ops += 20; return Cd(::pow(a.v, (double)b)); }
static inline Cd fabs(Cd a) { g_ops++; return Cd(::fabs(a.v)); }
static inline Cd floor(Cd a) { g_ops++; return Cd(::floor(a.v)); }
#define CD_SEL(f)                                                                 \
  static inline Cd f(Cd a, Cd b) { g_ops++; return Cd(::f(a.v, b.v)); }           \
  static inline Cd f(Cd a, double b) { g_ops++; return Cd(::f(a.v, b)); }         \
  static inline Cd f(double a, Cd b) { g_ops++; return Cd(::f(a, b.v)); }
CD_SEL(fmax) CD_SEL(fmin) CD_SEL(copysign)
#undef isnan
#undef isinf
static inline bool isnan(Cd a) { return a.v != a.v; }
static inline bool isinf(Cd a) { return ::fabs(a.v) > 1.7976931348623157e308; }

static void pfrx_oracle_ops_flush(void) {
  pthread_mutex_lock(&g_ops_mu);
  g_ops_total += g_ops;
  g_ops = 0;
  pthread_mutex_unlock(&g_ops_mu);
}

#define double Cd
#define PFRX_ORACLE_COUNTING 1
extern "C" {
#include "pfrx_oracle.c"
}
#undef double

extern "C" void pfrx_oracle_ops_reset(void) {
  pthread_mutex_lock(&g_ops_mu);
  g_ops_total = 0;
  pthread_mutex_unlock(&g_ops_mu);
  g_ops = 0;
}
// operations counted since the last reset, over all the threads of the pfrx_oracle_rstep calls (each
// worker flushes when it ends) plus the calling thread
extern "C" unsigned long long pfrx_oracle_ops_get(void) {
  pfrx_oracle_ops_flush();
  pthread_mutex_lock(&g_ops_mu);
  unsigned long long r = g_ops_total;
  pthread_mutex_unlock(&g_ops_mu);
  return r;
}
