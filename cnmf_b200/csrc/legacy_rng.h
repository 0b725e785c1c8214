// Host-side bit-exact reproduction of numpy's legacy RandomState(seed).standard_normal stream
// (MT19937 + polar Box-Muller "legacy gauss"), which is what scikit-learn's random NMF init
// draws from (sklearn _nmf.py:296-307 via check_random_state(int) -> np.random.RandomState).
#pragma once
#include <cstdint>

namespace cnmf {

struct LegacyRandomState {
  uint32_t key[624];
  int pos;
  int has_gauss;
  double gauss;
};

void legacy_seed(LegacyRandomState* st, uint32_t seed);
double legacy_double(LegacyRandomState* st);
double legacy_gauss(LegacyRandomState* st);

// One restart's random init, written straight into the packed device layout (host staging copy):
//   H  (k x n_features)  drawn FIRST, row-major            -> H[c * ldH + g]
//   W  (n_samples x k)   drawn second, row-major (j, c)    -> Wt[c * ldW + j]   (stored transposed)
// value = |avg * z| rounded to fp32 (the reference keeps fp64; see DESIGN.md precision note).
void nmf_random_init(uint32_t seed, double avg, int n_samples, int n_features, int k, float* Wt, long long ldW,
                     float* H, long long ldH);

}  // namespace cnmf
