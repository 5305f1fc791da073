// Host-only microbenchmark behind the staging pipeline's row copies (csrc/context.cu: copy_row): rows of a 4K float RGBA image
// (61440 bytes each) copied between two 133 MB buffers by T threads, with memcpy and with SSE2 streaming stores.
//   g++ -O2 -pthread row_copy.cpp -o /tmp/row_copy && /tmp/row_copy [threads]
#include <emmintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

static void copy_stream(char* dst, const char* src, size_t n)
{
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
}

int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 4, W = 3840, H = 2160;
    const size_t row = (size_t)W * 16, total = row * H;
    char* a = (char*)aligned_alloc(64, total);
    char* b = (char*)aligned_alloc(64, total);
    memset(a, 1, total);
    memset(b, 2, total);
    for (int mode = 0; mode < 2; mode++) {
        double best = 1e9;
        for (int rep = 0; rep < 7; rep++) {
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t]() {
                    for (int y = H * t / T; y < H * (t + 1) / T; y++) {
                        if (mode) copy_stream(b + (size_t)y * row, a + (size_t)y * row, row);
                        else memcpy(b + (size_t)y * row, a + (size_t)y * row, row);
                    }
                    _mm_sfence();
                });
            for (auto& x : th) x.join();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < best) best = ms;
        }
        printf("%-18s %d threads: %6.2f ms per 133 MB image  (%.1f GB/s copied)\n", mode ? "streaming stores" : "memcpy per row", T, best, total / best / 1e6);
    }
    return memcmp(a, b, total) != 0;
}
