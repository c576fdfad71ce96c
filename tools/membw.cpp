// Host memory read bandwidth with T threads over buffers of the bench's shape (64 x 5 MB):
// the ceiling of the FASTA clean+pack phase.  g++ -O3 -march=x86-64-v3 -pthread tools/membw.cpp -o /tmp/membw
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char **argv)
{
    const int nbuf = 64;
    const size_t len = 5062500;
    std::vector<std::vector<unsigned char>> bufs(nbuf);
    for (auto &b : bufs) { b.resize(len); memset(b.data(), 'A', len); }
    for (int T : {1, 4, 8, 16, 32}) {
        if (argc > 1 && atoi(argv[1]) != T) continue;
        double best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            std::atomic<int> next{0};
            std::atomic<long long> sink{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < T; t++)
                pool.emplace_back([&]() {
                    __m256i s = _mm256_setzero_si256();
                    for (;;) {
                        int i = next.fetch_add(1);
                        if (i >= nbuf) break;
                        const unsigned char *p = bufs[i].data();
                        for (size_t j = 0; j + 32 <= len; j += 32) s = _mm256_add_epi8(s, _mm256_loadu_si256((const __m256i *)(p + j)));
                    }
                    sink += _mm256_extract_epi8(s, 0);
                });
            for (auto &th : pool) th.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
        }
        printf("threads %2d: %.2f ms  %.1f GB/s\n", T, best * 1e3, nbuf * len / best / 1e9);
    }
}
