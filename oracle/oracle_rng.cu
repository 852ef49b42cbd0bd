// TEST INFRASTRUCTURE ONLY -- CPU oracle, RANSAC random streams.
// The reference seeds one cuRAND XORWOW generator per (superpixel, sample) with
// curand_init(1234, id, 0) (core/src/TPS_RGBD_kernels.cu:318-322) and draws with
// curand_uniform / curand (:347-348, :367).  cuRAND is a CUDA-toolkit dependency
// of the reference (curand_kernel.h, CUDA 12.9 here), not part of its tree, so the
// oracle runs the toolkit's own generator on the HOST: QUALIFIERS is the header's
// documented override point and the XORWOW code has host branches
// (curand_kernel.h:59-61, 607-611).  Compiled by nvcc as host code only.
#define QUALIFIERS static __forceinline__ __host__ __device__
#include <curand_kernel.h>
#include <vector>

struct OrcRng { std::vector<curandState> s; };

extern "C" OrcRng* orc_rng_create(int n, unsigned long long seed) {
  OrcRng* r = new OrcRng;
  r->s.resize(n);
  for (int i = 0; i < n; i++) curand_init(seed, (unsigned long long)i, 0ULL, &r->s[i]);
  return r;
}
extern "C" void orc_rng_destroy(OrcRng* r) { delete r; }
extern "C" unsigned int orc_rng_u32(OrcRng* r, int id) { return curand(&r->s[id]); }
extern "C" float orc_rng_uniform(OrcRng* r, int id) { return curand_uniform(&r->s[id]); }
// raw state words (v[5], d), for comparing against the device generator
extern "C" void orc_rng_state(const OrcRng* r, int id, unsigned int* out6) {
  const curandState& s = r->s[id];
  out6[0] = s.d;
  for (int k = 0; k < 5; k++) out6[1 + k] = s.v[k];
}
