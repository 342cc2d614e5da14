// host_pack_bench.cpp -- how fast can the host cores cut 168-byte mp_float_t records down to the 40 bytes the device needs of A and B
// (first four residues, sign, exponent, upper interval bound)?  Decides whether a lean host->device format pays.
// g++ -O2 -pthread tools/host_pack_bench.cpp -o tools/host_pack_bench
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char **argv) {
    const size_t n = argc > 1 ? atoll(argv[1]) : (size_t) 1 << 24;
    const size_t rs = 168, ls = 40;
    char *src = (char *) aligned_alloc(4096, n * rs), *dst = (char *) aligned_alloc(4096, n * ls);
    memset(src, 1, n * rs); memset(dst, 0, n * ls);
    for (int T : {1, 4, 8, 16, 32}) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([=]() {
                const size_t b = n * t / T, e = n * (t + 1) / T;
                for (size_t i = b; i < e; ++i) {
                    const char *r = src + i * rs;
                    char *o = dst + i * ls;
                    memcpy(o, r, 16);                 // digits 0..3
                    memcpy(o + 16, r + 128, 8);       // sign, exp
                    memcpy(o + 24, r + 152, 16);      // eval[1]
                }
            });
        for (auto &x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %2d: %.1f ms, %.1f GB/s of source records, %.1f GB/s packed (hw threads %u)\n", T, s * 1e3, n * rs / s / 1e9, n * ls / s / 1e9, std::thread::hardware_concurrency());
    }
    return 0;
}
