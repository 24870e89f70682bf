// pack_probe.cpp -- how fast can the host's cores pack 10-bit samples (uint16 containers) three to a 32-bit word?  Decides whether a packed
// upload (11.1 MB instead of 16.6 MB per 2160p picture) would pay on a host link that is the end-to-end limit.
// build: g++ -O3 -march=native -std=c++17 -pthread tools/pack_probe.cpp -o tools/pack_probe_bin
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void pack_rows(const uint16_t* src, size_t stride, uint32_t* dst, size_t dstWords, int width, int r0, int r1)
{
    for (int r = r0; r < r1; r++)
    {
        const uint16_t* s = src + (size_t)r * stride;
        uint32_t* d = dst + (size_t)r * dstWords;
        for (int i = 0; i < width / 3; i++) d[i] = (uint32_t)s[3 * i] | ((uint32_t)s[3 * i + 1] << 10) | ((uint32_t)s[3 * i + 2] << 20);
    }
}

int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 16, W = 3840, H = 2160, NP = 16;
    const size_t stride = 4032, words = W / 3;
    std::vector<uint16_t*> src(NP); std::vector<uint32_t*> dst(NP);
    for (int p = 0; p < NP; p++)
    {
        src[p] = (uint16_t*)aligned_alloc(4096, stride * H * 2); dst[p] = (uint32_t*)aligned_alloc(4096, words * H * 4);
        for (size_t i = 0; i < stride * H; i++) src[p][i] = (uint16_t)((i * 2654435761u >> 12) & 1023);
        memset(dst[p], 0, words * H * 4);
    }
    for (int rep = 0; rep < 3; rep++)
    {
        auto t0 = std::chrono::steady_clock::now();
        for (int p = 0; p < NP; p++)
        {   // one picture at a time, its rows split over T threads (what an upload call would do)
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back(pack_rows, src[p], stride, dst[p], words, W, H * t / T, H * (t + 1) / T);
            for (auto& x : th) x.join();
        }
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.3f ms per picture (incl. thread create/join), %.1f GB/s of picture bytes\n", T, dt / NP * 1e3, (double)W * H * 2 * NP / dt / 1e9);
    }
    return 0;
}
