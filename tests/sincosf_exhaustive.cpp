// Host build of the product's glibc sinf/cosf replica, compared with libm for every float in [lo, hi].
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "glibc_sincosf.h"

int main(int argc, char** argv) {
    float lo = argc > 1 ? (float)atof(argv[1]) : 0.f, hi = argc > 2 ? (float)atof(argv[2]) : 6.2832f;
    uint32_t ulo, uhi;
    memcpy(&ulo, &lo, 4);
    memcpy(&uhi, &hi, 4);
    long bad = 0;
    for (uint32_t u = ulo; u <= uhi; ++u) {
        float f, s, c;
        memcpy(&f, &u, 4);
        orbb::sincosf_glibc(f, &s, &c);
        const float rs = sinf(f), rc = cosf(f);
        if (memcmp(&s, &rs, 4) || memcmp(&c, &rc, 4)) {
            if (bad < 5) printf("mismatch at %a: sin %a vs %a, cos %a vs %a\n", f, s, rs, c, rc);
            ++bad;
        }
    }
    printf("checked %u mismatches %ld\n", uhi - ulo + 1, bad);
    return bad ? 1 : 0;
}
