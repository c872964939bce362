// End-to-end timing of the drop-in boundary itself: gl_commit (include/gl_commit.h) called from C++ with pinned HOST columns, no
// Python in the loop — what a compiled host (the Rust shim of INTEGRATION.md, include/gl_plonky2.hpp) pays per
// PolynomialBatch::from_values.  Two modes per shape:
//   cap        outputs stay device-resident behind the handle, only the cap comes back (the intended integration)
//   copyback   coefficients, leaves and digests are also copied to pinned host buffers (a host that materialises
//              MerkleTree::leaves / digests, SURVEY.md §8d "what the Rust drop-in would pay")
// Wall-clock around the synchronous calls (std::chrono), W warm-up + K timed calls, one JSON line per (shape, mode).
// Build: g++ -std=c++17 -O2 -I include tools/cbench.cpp -o build/cbench -L plonky2.5_b200 -lgl_commit -pthread
// Usage: cbench [log_n n_cols rate_bits cap_height mode(0 cap / 1 copyback pinned / 2 copyback pageable)]...   (default: cfg2 x3 modes, cfg3 cap)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "gl_commit.h"

static void fill(uint64_t* col, uint64_t n, uint64_t first_index) {   // tests/oracle_c.py · splitmix_columns (seed 2025)
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t z = (0x706C6F6E6B7932ULL ^ 2025) + (first_index + i + 1) * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        col[i] = z >= P ? z - P : z;
    }
}

static int run(gl_ctx* ctx, unsigned log_n, unsigned n_cols, unsigned r, unsigned h, int mode, int warmup, int steps) {
    const bool copyback = mode != 0, pageable = mode == 2;   // 2: outputs in plain malloc memory (a Rust Vec / numpy array)
    const uint64_t N = 1ULL << log_n, R = N << r, n_dig = 2 * (R - (1ULL << h));
    uint64_t* in = (uint64_t*)gl_host_alloc(N * n_cols * 8);
    if (!in) { std::printf("{\"error\": \"pinned allocation failed\"}\n"); return 1; }
    {
        std::vector<std::thread> th;
        const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([=] { for (unsigned j = t; j < n_cols; j += T) fill(in + (uint64_t)j * N, N, (uint64_t)j * N); });
        for (auto& x : th) x.join();
    }
    std::vector<const uint64_t*> cols(n_cols);
    for (unsigned j = 0; j < n_cols; j++) cols[j] = in + (uint64_t)j * N;
    uint64_t *oc = nullptr, *ol = nullptr, *od = nullptr;
    if (copyback) {
        oc = (uint64_t*)(pageable ? std::malloc(N * n_cols * 8) : gl_host_alloc(N * n_cols * 8));
        ol = (uint64_t*)(pageable ? std::malloc(R * n_cols * 8) : gl_host_alloc(R * n_cols * 8));
        od = (uint64_t*)(pageable ? std::malloc(n_dig * 32) : gl_host_alloc(n_dig * 32));
        if (!oc || !ol || !od) { std::printf("{\"error\": \"pinned allocation failed\"}\n"); return 1; }
    }
    std::vector<uint64_t> cap(4ULL << h), cap0;
    double total_ms = 0, best_ms = 1e30;
    float st[GL_N_STAGES] = {};
    uint32_t ln[GL_N_STAGES] = {};
    for (int it = 0; it < warmup + steps; it++) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = gl_commit(ctx, cols.data(), n_cols, log_n, r, h, 0, oc, ol, od, cap.data(), nullptr);
        auto t1 = std::chrono::steady_clock::now();
        if (rc != GL_OK) { std::printf("{\"error\": \"gl_commit: %s: %s\"}\n", gl_strerror(rc), gl_ctx_last_error(ctx)); return 1; }
        if (it == 0) cap0 = cap;
        if (cap != cap0) { std::printf("{\"error\": \"cap changed between identical commits\"}\n"); return 1; }
        if (it >= warmup) {
            const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            total_ms += ms;
            best_ms = std::min(best_ms, ms);
        }
    }
    gl_ctx_stage_times(ctx, st, ln);
    const char* verified = "n/a";
    if (copyback) {   // the copies that overlapped the hashing must equal what the resident batch holds
        gl_handle hd = 0;
        std::memset(oc, 0xA5, N * n_cols * 8); std::memset(ol, 0xA5, R * n_cols * 8); std::memset(od, 0xA5, n_dig * 32);
        int rc = gl_commit(ctx, cols.data(), n_cols, log_n, r, h, 0, oc, ol, od, cap.data(), &hd);
        std::vector<uint64_t> a(N * n_cols), b(R * n_cols), d(n_dig * 4);
        if (rc == GL_OK) rc = gl_tree_read(ctx, hd, GL_PART_COEFFS, a.data());
        if (rc == GL_OK) rc = gl_tree_read(ctx, hd, GL_PART_LEAVES, b.data());
        if (rc == GL_OK) rc = gl_tree_read(ctx, hd, GL_PART_DIGESTS, d.data());
        if (rc != GL_OK) { std::printf("{\"error\": \"verification: %s: %s\"}\n", gl_strerror(rc), gl_ctx_last_error(ctx)); return 1; }
        const bool ok = !std::memcmp(a.data(), oc, a.size() * 8) && !std::memcmp(b.data(), ol, b.size() * 8) && !std::memcmp(d.data(), od, d.size() * 8);
        gl_tree_free(ctx, hd);
        if (!ok) { std::printf("{\"error\": \"copy-back differs from the resident batch\"}\n"); return 1; }
        verified = "copy-back == gl_tree_read of the resident batch";
    }
    const double ms = total_ms / steps;
    uint32_t launches = 0;
    for (auto x : ln) launches += x;
    std::printf("{\"metric\": \"commit Melem/s, end to end through gl_commit from C++ (host columns in)\", \"mode\": \"%s\", \"log_n\": %u, \"n_cols\": %u, "
                "\"rate_bits\": %u, \"cap_height\": %u, \"value\": %.1f, \"unit\": \"Melem/s\", \"ms_per_call\": %.3f, \"best_ms\": %.3f, \"steps\": %d, \"warmup\": %d, "
                "\"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"gpu_launches_per_call\": %u, "
                "\"stage_ms_last\": {\"h2d\": %.3f, \"transpose\": %.3f, \"intt\": %.3f, \"lde\": %.3f, \"leaf_hash\": %.3f, \"tree\": %.3f, \"d2h\": %.3f}, "
                "\"cap0\": \"%016llx\", \"verified\": \"%s\"}\n",
                pageable ? "copyback (coeffs + leaves + digests + cap to PAGEABLE host memory)" : copyback ? "copyback (coeffs + leaves + digests + cap to pinned host)" : "cap (outputs stay in HBM)", log_n, n_cols, r, h,
                (double)N * n_cols / (ms * 1e3), ms, best_ms, steps, warmup, (unsigned long long)(N * n_cols * 8),
                (unsigned long long)((copyback ? N * n_cols * 8 + R * n_cols * 8 + n_dig * 32 : 0) + (32ULL << h)), launches, st[0], st[1], st[2], st[3], st[4],
                st[5], st[6], (unsigned long long)cap[0], verified);
    std::fflush(stdout);
    gl_host_free(in);
    if (copyback && pageable) { std::free(oc); std::free(ol); std::free(od); }
    else if (copyback) { gl_host_free(oc); gl_host_free(ol); gl_host_free(od); }
    return 0;
}

int main(int argc, char** argv) {
    gl_ctx* ctx = nullptr;
    int rc = gl_ctx_create(&ctx, 0);
    if (rc != GL_OK) { std::printf("{\"error\": \"gl_ctx_create: %s (no CPU fallback)\"}\n", gl_strerror(rc)); return 2; }
    int bad = 0;
    if (argc >= 6) {
        for (int i = 1; i + 4 < argc; i += 5)
            bad |= run(ctx, atoi(argv[i]), atoi(argv[i + 1]), atoi(argv[i + 2]), atoi(argv[i + 3]), atoi(argv[i + 4]), 3, 5);
    } else {
        bad |= run(ctx, 16, 135, 3, 4, 0, 3, 10);
        bad |= run(ctx, 16, 135, 3, 4, 1, 3, 10);
        bad |= run(ctx, 16, 135, 3, 4, 2, 3, 10);
        bad |= run(ctx, 20, 135, 3, 4, 0, 3, 5);
    }
    gl_ctx_destroy(ctx);
    return bad;
}
