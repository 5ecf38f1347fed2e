// Host build of scat_b200/csrc/eval_math.cuh for the CPU test suite (tests/test_eval_metrics_host.py):
// reads B, n and two [B,n,3] fp32 arrays from stdin (binary), writes the aligned points and scales to stdout.
#include <stdio.h>
#include <stdlib.h>

#include "../../scat_b200/csrc/eval_math.cuh"

int main() {
    int hdr[2];
    if (fread(hdr, sizeof(int), 2, stdin) != 2) return 2;
    const int B = hdr[0], n = hdr[1];
    const size_t m = (size_t)B * n * 3;
    float* s1 = (float*)malloc(m * 4);
    float* s2 = (float*)malloc(m * 4);
    float* out = (float*)malloc(m * 4);
    float* sc = (float*)malloc((size_t)B * 4);
    if (fread(s1, 4, m, stdin) != m || fread(s2, 4, m, stdin) != m) return 3;
    for (int b = 0; b < B; ++b) scat::evalm::similarity_align(s1 + (size_t)b * n * 3, s2 + (size_t)b * n * 3, n, out + (size_t)b * n * 3, sc + b);
    fwrite(out, 4, m, stdout);
    fwrite(sc, 4, B, stdout);
    return 0;
}
