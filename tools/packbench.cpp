// Pack-phase ceiling without CUDA, Python or pinned memory: T threads clean + pack 64 synthetic 5 Mbp FASTA
// texts (80-column lines) held in malloc'ed memory into malloc'ed output.
// g++ -O3 -march=x86-64-v3 -std=c++17 -pthread -Isupersampler_b200/csrc/host -Iinclude tools/packbench.cpp \
//     supersampler_b200/csrc/host/seqio.cpp -Lsupersampler_b200/lib -lspsp_b200 -lz -Wl,-rpath,$PWD/supersampler_b200/lib -o /tmp/packbench
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>
#include "seqio.h"
using namespace spsp_host;
int main(int argc, char **argv)
{
    const int nbuf = 64;
    const size_t nbases = 5000000;
    std::vector<std::vector<uint8_t>> fa(nbuf);
    std::mt19937_64 rng(1);
    for (auto &f : fa) {
        f.reserve(nbases + nbases / 80 + 16);
        f.push_back('>'); f.push_back('g'); f.push_back('\n');
        for (size_t i = 0; i < nbases; i++) { f.push_back("ACGT"[rng() & 3]); if (i % 80 == 79) f.push_back('\n'); }
    }
    for (int T : {1, 4, 8, 16}) {
        if (argc > 1 && atoi(argv[1]) != T) continue;
        std::vector<std::unique_ptr<PackedInput>> out;
        for (int t = 0; t < T; t++) { out.emplace_back(new PackedInput(false)); out.back()->words.reserve(nbases / 16 + 1024); }
        double best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            std::atomic<int> next{0};
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < T; t++)
                pool.emplace_back([&, t]() {
                    for (;;) {
                        int i = next.fetch_add(1);
                        if (i >= nbuf) break;
                        pack_fasta_buffer(fa[i].data(), fa[i].size(), 31, *out[t]);
                    }
                });
            for (auto &th : pool) th.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
        }
        printf("threads %2d: %.2f ms  %.1f GB/s of text (%.2f GB/s per thread)\n", T, best * 1e3, nbuf * fa[0].size() / best / 1e9,
               nbuf * fa[0].size() / best / 1e9 / T);
    }
}
