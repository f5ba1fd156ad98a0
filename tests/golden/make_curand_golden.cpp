// Generates tests/golden/curand_philox.txt from the CUDA toolkit's own host implementation of
// curandStatePhilox4_32_10 (curand_kernel.h, the third-party dependency the reference draws every
// random number from: qudarap/qrng.h:124-148, src/torch.cpp:10-11).
//   /usr/bin/g++ -x c++ -I/usr/local/cuda/include make_curand_golden.cpp -o /tmp/mkgold && /tmp/mkgold > curand_philox.txt
#include <curand_kernel.h>
#include <cstdio>
int main() {
    struct Case { unsigned long long seed, sub, off, skip; int n; };
    Case cases[] = {{0, 0, 0, 0, 16}, {0, 1, 0, 0, 16}, {0, 9, 0, 0, 8}, {0, 999999, 0, 0, 8}, {0, 5000000000ull, 0, 0, 8},
                    {42, 12345678901ull, 7, 300000, 12}, {0, 17, 0, 100000, 8}, {0, 17, 0, 200000, 8}, {7, 3, 2, 1, 9}, {0xffffffffffull, 2, 3, 5, 9}};
    for (auto& c : cases) {
        curandStatePhilox4_32_10 r;
        curand_init(c.seed, c.sub, c.off, &r);
        skipahead(c.skip, &r);
        printf("%llu %llu %llu %llu %d", c.seed, c.sub, c.off, c.skip, c.n);
        for (int i = 0; i < c.n; i++) { float u = curand_uniform(&r); unsigned b; __builtin_memcpy(&b, &u, 4); printf(" %u", b); }
        printf("\n");
    }
    return 0;
}
