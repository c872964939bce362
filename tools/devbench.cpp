// Kernel-iteration bench without Python: one device-resident commit (gl_dev_commit — the call bench.py times as `value`) of the
// SplitMix64 seed-1 input (== bench.py · synth_columns / tests/oracle_c.py · splitmix_columns), W warm-up + K timed calls, best and
// mean stage times (CUDA events inside the library) and the first cap words as one JSON line.  Starts in milliseconds, so it fits
// short GPU calls and ncu captures (`ncu -k regex:ntt_pass ... tests/cpp/build/devbench 20 135 3 4 1 1`).
// Build: g++ -std=c++17 -O2 -I include tools/devbench.cpp -o tests/cpp/build/devbench -L plonky2.5_b200 -lgl_commit -pthread
// Usage: devbench [log_n n_cols rate_bits cap_height warmup steps]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "gl_commit.h"

static void fill(uint64_t* col, uint64_t n, uint64_t first_index) {
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t z = (0x706C6F6E6B7932ULL ^ 1) + (first_index + i + 1) * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        col[i] = z >= P ? z - P : z;
    }
}

int main(int argc, char** argv) {
    const unsigned log_n = argc > 1 ? atoi(argv[1]) : 20, n_cols = argc > 2 ? atoi(argv[2]) : 135, r = argc > 3 ? atoi(argv[3]) : 3,
                   h = argc > 4 ? atoi(argv[4]) : 4;
    const int warmup = argc > 5 ? atoi(argv[5]) : 3, steps = argc > 6 ? atoi(argv[6]) : 5;
    const uint64_t N = 1ULL << log_n;
    gl_ctx* ctx = nullptr;
    if (int rc = gl_ctx_create(&ctx, 0)) { std::printf("{\"error\": \"gl_ctx_create: %s\"}\n", gl_strerror(rc)); return 1; }
    std::vector<uint64_t> host(N * n_cols);
    {
        std::vector<std::thread> th;
        const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t] { for (unsigned j = t; j < n_cols; j += T) fill(host.data() + (uint64_t)j * N, N, (uint64_t)j * N); });
        for (auto& x : th) x.join();
    }
    uint64_t* d = nullptr;
    if (gl_dev_alloc(ctx, N * n_cols, &d) || gl_dev_upload(ctx, host.data(), d, N * n_cols)) {
        std::printf("{\"error\": \"%s\"}\n", gl_ctx_last_error(ctx));
        return 1;
    }
    std::vector<uint64_t> cap(4ULL << h);
    float best[GL_N_STAGES], sum[GL_N_STAGES] = {}, ms[GL_N_STAGES];
    std::fill(best, best + GL_N_STAGES, 1e30f);
    uint32_t launches[GL_N_STAGES] = {};
    for (int it = 0; it < warmup + steps; it++) {
        gl_handle hd = 0;
        if (gl_dev_commit(ctx, d, N, n_cols, log_n, r, h, 0, cap.data(), &hd)) { std::printf("{\"error\": \"%s\"}\n", gl_ctx_last_error(ctx)); return 1; }
        gl_tree_free(ctx, hd);
        if (it < warmup) continue;
        gl_ctx_stage_times(ctx, ms, launches);
        for (int i = 0; i < GL_N_STAGES; i++) { best[i] = std::min(best[i], ms[i]); sum[i] += ms[i]; }
    }
    static const char* names[GL_N_STAGES] = {"h2d", "transpose", "intt", "lde", "leaf_hash", "tree", "d2h"};
    float tot = 0;
    for (int i = 0; i < GL_N_STAGES; i++) tot += sum[i] / steps;
    std::printf("{\"shape\": [%u, %u, %u, %u], \"melem_per_s\": %.1f, \"ms\": %.3f, \"mean_ms\": {", log_n, n_cols, r, h, N * n_cols / (tot * 1e-3) / 1e6, tot);
    for (int i = 0; i < GL_N_STAGES; i++) std::printf("%s\"%s\": %.3f", i ? ", " : "", names[i], sum[i] / steps);
    std::printf("}, \"best_ms\": {");
    for (int i = 0; i < GL_N_STAGES; i++) std::printf("%s\"%s\": %.3f", i ? ", " : "", names[i], best[i]);
    std::printf("}, \"cap0\": [%llu, %llu]}\n", (unsigned long long)cap[0], (unsigned long long)cap[1]);
    gl_dev_free(ctx, d);
    gl_ctx_destroy(ctx);
    return 0;
}
