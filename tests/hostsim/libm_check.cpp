// libm_check — TEST INFRASTRUCTURE.  Compares csrc/libm_exact.cuh (compiled for the host) with this
// machine's glibc over EVERY float of the argument ranges the render path uses:
//   sinf, cosf on [0, 8)   logf on (0, 1]   expf on [-87, 4]   log2f on (0, 1e30]   powf(x, {1, 0.8, 1.35, 2.2}) on [-16, 16]
// Prints one line per function: values tested, mismatches.   g++ -O2 -ffp-contract=off -DYB_HOSTSIM
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../../yart_b200/csrc/libm_exact.cuh"

template <class F, class G>
static void sweep(const char* name, uint32_t lo, uint32_t hi, bool negative, F mine, G ref, uint32_t stride) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> bad{0}, total{0};
  std::atomic<uint32_t> firstBad{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      uint64_t b = 0, n = 0;
      for (uint64_t u = uint64_t(lo) + uint64_t(t) * stride; u <= hi; u += uint64_t(nt) * stride) {
        uint32_t bits = uint32_t(u) | (negative ? 0x80000000u : 0u);
        float x;
        memcpy(&x, &bits, 4);
        float a = mine(x), r = ref(x);
        uint32_t ab, rb;
        memcpy(&ab, &a, 4);
        memcpy(&rb, &r, 4);
        n++;
        if (ab != rb && !(a != a && r != r)) {
          if (b == 0) firstBad = bits;
          b++;
        }
      }
      bad += b;
      total += n;
    });
  for (auto& x : th) x.join();
  uint32_t fb = firstBad;
  float fx;
  memcpy(&fx, &fb, 4);
  printf("%s %s: tested %llu mismatches %llu", name, negative ? "(negative)" : "", (unsigned long long)total.load(),
         (unsigned long long)bad.load());
  if (bad) printf(" first x=%a mine=%a glibc=%a", fx, mine(fx), ref(fx));
  printf("\n");
}

static uint32_t bitsOf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

int main(int argc, char** argv) {
  uint32_t stride = argc > 1 ? uint32_t(atoi(argv[1])) : 1;
  sweep("sinf", 0, bitsOf(8.0f), false, yb::sinfExact, [](float x) { return sinf(x); }, stride);
  sweep("cosf", 0, bitsOf(8.0f), false, yb::cosfExact, [](float x) { return cosf(x); }, stride);
  sweep("logf", 1, bitsOf(1.0f), false, yb::logfExact, [](float x) { return logf(x); }, stride);
  sweep("expf", 0, bitsOf(4.0f), false, yb::expfExact, [](float x) { return expf(x); }, stride);
  sweep("expf", 0, bitsOf(87.0f), true, yb::expfExact, [](float x) { return expf(x); }, stride);
  sweep("log2f", 1, bitsOf(1e30f), false, yb::log2fExact, [](float x) { return log2f(x); }, stride);
  const float ys[4] = {1.0f, 0.8f, 1.35f, 2.2f};  // AgX look powers and the final 2.2
  for (float y : ys) {
    char name[32];
    snprintf(name, sizeof name, "powf(x,%g)", y);
    sweep(name, 0, bitsOf(16.0f), false, [y](float x) { return yb::powfExact(x, y); }, [y](float x) { return powf(x, y); }, stride);
    sweep(name, 1, bitsOf(16.0f), true, [y](float x) { return yb::powfExact(x, y); }, [y](float x) { return powf(x, y); }, stride);
  }
  return 0;
}
