/* Independent Philox4x32-10: NVIDIA's cuRAND header compiled for the host (test infrastructure only).
 * Prints 4 output words for each (ctr[4], key[2]) read from stdin as 6 hex words per line. */
#include <cstdint>
#include <cstdio>
#define QUALIFIERS static inline
#include <vector_types.h>
#include <curand_philox4x32_x.h>
int main() {
  unsigned c0, c1, c2, c3, k0, k1;
  while (scanf("%x %x %x %x %x %x", &c0, &c1, &c2, &c3, &k0, &k1) == 6) {
    uint4 c = {c0, c1, c2, c3};
    uint2 k = {k0, k1};
    uint4 o = curand_Philox4x32_10(c, k);
    printf("%08x %08x %08x %08x\n", o.x, o.y, o.z, o.w);
  }
  return 0;
}
