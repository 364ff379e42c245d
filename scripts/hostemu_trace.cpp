// Analysis tool: the test harness's host build with every step attempt of one chosen reach printed (SP_ATTEMPT_HOOK).
//   g++ -O2 -std=c++17 -shared -fPIC -o build/libhostemu_trace.so scripts/hostemu_trace.cpp ; TRACE_S=2 TRACE_D0=20 TRACE_D1=22
#include <cstdio>
#include <cstdlib>
static int trace_env(const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; }
static void trace_hook(int s, int day, int k, double t, double hh, double en2, bool acc) {
  static int S = trace_env("TRACE_S", -1), d0 = trace_env("TRACE_D0", 0), d1 = trace_env("TRACE_D1", 0);
  if (s == S && day >= d0 && day < d1) fprintf(stderr, "s %d day %d k %2d t %.6f h %.3e en2 %.3e %s\n", s, day, k, t, hh, en2, acc ? "" : "REJ");
}
#define SP_ATTEMPT_HOOK(io, day, k, t, hh, en2, acc) trace_hook((io).s, day, k, t, hh, en2, acc)
#include "../tests/hostemu/hostemu.cpp"
