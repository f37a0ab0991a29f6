// host memcpy scaling on the GPU box (pageable -> page-locked-like buffer), used to size par_memcpy's thread count
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main() {
    const size_t n = (size_t)1 << 30;
    char* a = (char*)malloc(n);
    char* b = (char*)malloc(n);
    memset(a, 1, n);
    memset(b, 2, n);
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    for (int t : {1, 2, 4, 6, 8, 12, 16, 24, 32}) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            const size_t per = n / t;
            for (int i = 0; i < t; ++i) th.emplace_back([=] { memcpy(b + i * per, a + i * per, per); });
            for (auto& x : th) x.join();
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        printf("threads %2d: %.1f GB/s\n", t, n / best / 1e9);
    }
    return 0;
}
