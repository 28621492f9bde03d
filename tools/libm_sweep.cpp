#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
// TEST INFRASTRUCTURE: exhaustive comparison of the restated libm functions of bbm_b200/csrc/bbmcu_libm.cuh (compiled for the
// host) with this machine's glibc, argument by argument over whole float ranges, both signs.
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -Ibbm_b200/csrc -o /tmp/libm_sweep tools/libm_sweep.cpp -lpthread
// Last run (glibc 2.39, x86-64, 8 threads, 13 s): tanf [0, 100] 0 mismatches of 2.2e9; erfcf all floats 0 of 4.3e9.
#include "bbmcu_libm.cuh"
#ifdef LOGSWEEP
#include "bbmcu_math.cuh"
#endif
using namespace bbmcu;
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
template<class F, class G> void sweep(const char* name, uint32_t lo, uint32_t hi, F mine, G ref)
{
  std::atomic<uint64_t> bad{0}; std::atomic<uint32_t> first_bad{0};
  int T = 8; std::vector<std::thread> th;
  for(int t=0; t < T; ++t) th.emplace_back([&, t] {
    uint64_t b = 0;
    for(uint64_t u = (uint64_t)lo + t; u <= hi; u += T) {
      float x; uint32_t uu = (uint32_t)u; memcpy(&x, &uu, 4);
      for(int sgn = 0; sgn < 2; ++sgn) {
        float xx = sgn ? -x : x;
        float a = mine(xx), r = ref(xx);
        if(a != a && r != r) continue;
        if(bits(a) != bits(r)) { if(!b && !bad) first_bad = bits(xx); ++b; }
      }
    }
    bad += b; });
  for(auto& x : th) x.join();
  float fb; uint32_t f = first_bad; memcpy(&fb, &f, 4);
  printf("%s [%08x, %08x] both signs: %llu mismatches", name, lo, hi, (unsigned long long)bad.load());
  if(bad) printf("  first at x = %.9g (%08x): mine %.9g ref %.9g", fb, f, mine(fb), ref(fb));
  printf("\n");
}
int main()
{
#ifdef LOGSWEEP
  // w = (float)(-log((1 - a)(1 + a))) of erfinv (bbmcu_math.cuh) with fast_log_pos against the library's double log, every float a in [0, 1]
  //   g++ -O2 -std=c++17 -ffp-contract=off -DLOGSWEEP -I/usr/local/cuda/include -Ibbm_b200/csrc -o /tmp/log_sweep tools/libm_sweep.cpp -lpthread
  // Last run (glibc 2.39, 9 s): 0 mismatches of 2 x 1 065 353 217 arguments
  sweep("erfinv w", 0x00000000, 0x3f800000, [](float a) { return (float)(-fast_log_pos((1.0 - (double)a) * (1.0 + (double)a))); },
        [](float a) { return (float)(-log((1.0 - (double)a) * (1.0 + (double)a))); });
  return 0;
#endif
  sweep("tanf ", 0x00000000, 0x42c80000, [](float x) { return glibc_tanf(x); }, [](float x) { return tanf(x); });       // |x| <= 100
  sweep("erfcf", 0x00000000, 0x7f800000, [](float x) { return glibc_erfcf(x); }, [](float x) { return erfcf(x); });
  return 0;
}
