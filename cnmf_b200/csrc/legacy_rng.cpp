#include "legacy_rng.h"

#include <cmath>

namespace cnmf {

namespace {
constexpr int RK_N = 624, RK_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;

inline void mt_gen(LegacyRandomState* st) {
  uint32_t* key = st->key;
  uint32_t y;
  int i;
  for (i = 0; i < RK_N - RK_M; i++) {
    y = (key[i] & UPPER) | (key[i + 1] & LOWER);
    key[i] = key[i + RK_M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  for (; i < RK_N - 1; i++) {
    y = (key[i] & UPPER) | (key[i + 1] & LOWER);
    key[i] = key[i + (RK_M - RK_N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  y = (key[RK_N - 1] & UPPER) | (key[0] & LOWER);
  key[RK_N - 1] = key[RK_M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  st->pos = 0;
}

inline uint32_t mt_next(LegacyRandomState* st) {
  if (st->pos == RK_N) mt_gen(st);
  uint32_t y = st->key[st->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
}  // namespace

void legacy_seed(LegacyRandomState* st, uint32_t seed) {
  for (int pos = 0; pos < RK_N; pos++) {
    st->key[pos] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + static_cast<uint32_t>(pos) + 1u;
  }
  st->pos = RK_N;
  st->has_gauss = 0;
  st->gauss = 0.0;
}

double legacy_double(LegacyRandomState* st) {
  const int32_t a = static_cast<int32_t>(mt_next(st) >> 5), b = static_cast<int32_t>(mt_next(st) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

double legacy_gauss(LegacyRandomState* st) {
  if (st->has_gauss) {
    const double t = st->gauss;
    st->has_gauss = 0;
    st->gauss = 0.0;
    return t;
  }
  double f, x1, x2, r2;
  do {
    x1 = 2.0 * legacy_double(st) - 1.0;
    x2 = 2.0 * legacy_double(st) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  f = std::sqrt(-2.0 * std::log(r2) / r2);
  st->gauss = f * x1;
  st->has_gauss = 1;
  return f * x2;
}

void nmf_random_init(uint32_t seed, double avg, int n_samples, int n_features, int k, float* Wt, long long ldW,
                     float* H, long long ldH) {
  LegacyRandomState st;
  legacy_seed(&st, seed);
  for (int c = 0; c < k; ++c)
    for (int g = 0; g < n_features; ++g) H[c * ldH + g] = static_cast<float>(std::fabs(avg * legacy_gauss(&st)));
  for (int j = 0; j < n_samples; ++j)
    for (int c = 0; c < k; ++c) Wt[c * ldW + j] = static_cast<float>(std::fabs(avg * legacy_gauss(&st)));
}

}  // namespace cnmf
