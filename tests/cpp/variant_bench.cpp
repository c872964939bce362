// Development aid: check and time compile-time variants of libgl_commit.so in ONE short GPU call, without Python.
//     variant_bench [--log-n 20] [--cols 135] [--rate-bits 3] lib1.so lib2.so ...
// For every library (one forked child each, so every variant gets a fresh CUDA runtime):
//   1. parity: a 2^10 x 135, rate_bits 3 commit compared word for word with the C oracle (coefficients, leaves, digests, cap)
//      and a 2^7 x 20 commit with non-canonical inputs;
//   2. timing: W=2 + K=4 gl_commit calls at the bench shape from pinned host columns, best stage times (CUDA events inside the
//      library) and the first cap word, which must agree between all variants (the first library is the reference build).
// One JSON line per library.  The oracle is linked as the checker only (this file lives under tests/).
// tests/variants.py builds the variants (python tests/variants.py build) and is the torch-side twin of this tool.
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gl_commit.h"

extern "C" int glo_commit(const uint64_t* const* cols, uint32_t n_cols, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int is_coeffs,
                          uint64_t* out_coeffs, uint64_t* out_leaves, uint64_t* out_digests, uint64_t* out_cap, double* stage_s);

namespace {
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;

void fill(uint64_t* col, uint64_t n, uint64_t first_index, uint64_t seed, bool canonical) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t z = (0x706C6F6E6B7932ULL ^ seed) + (first_index + i + 1) * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        col[i] = (canonical && z >= P) ? z - P : z;
    }
}

struct Lib {
    void* h = nullptr;
    decltype(&gl_ctx_create) ctx_create;
    decltype(&gl_ctx_destroy) ctx_destroy;
    decltype(&gl_ctx_last_error) last_error;
    decltype(&gl_commit) commit;
    decltype(&gl_tree_read) tree_read;
    decltype(&gl_tree_free) tree_free;
    decltype(&gl_ctx_stage_times) stage_times;
    decltype(&gl_host_alloc) host_alloc;
    bool open(const char* path) {
        h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) { std::printf("{\"lib\": \"%s\", \"error\": \"dlopen: %s\"}\n", path, dlerror()); return false; }
#define SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(h, name)); if (!field) { std::printf("{\"lib\": \"%s\", \"error\": \"missing %s\"}\n", path, name); return false; }
        SYM(ctx_create, "gl_ctx_create") SYM(ctx_destroy, "gl_ctx_destroy") SYM(last_error, "gl_ctx_last_error") SYM(commit, "gl_commit")
        SYM(tree_read, "gl_tree_read") SYM(tree_free, "gl_tree_free") SYM(stage_times, "gl_ctx_stage_times") SYM(host_alloc, "gl_host_alloc")
#undef SYM
        return true;
    }
};

bool parity(Lib& L, gl_ctx* ctx, unsigned log_n, unsigned c, unsigned r, unsigned h, bool canonical, uint64_t seed) {
    const uint64_t n = 1ULL << log_n, R = n << r, nd = 8 * (R - (1ULL << h));
    std::vector<uint64_t> in(n * c);
    fill(in.data(), n * c, 0, seed, canonical);
    std::vector<const uint64_t*> cols(c);
    for (unsigned j = 0; j < c; j++) cols[j] = in.data() + (uint64_t)j * n;
    std::vector<uint64_t> rc_(n * c), rl(R * c), rd(nd + 4), rcap(4ULL << h), gc(n * c), gl(R * c), gd(nd + 4), gcap(4ULL << h);
    if (glo_commit(cols.data(), c, log_n, r, h, 0, rc_.data(), rl.data(), rd.data(), rcap.data(), nullptr)) return false;
    gl_handle hd = 0;
    if (L.commit(ctx, cols.data(), c, log_n, r, h, 0, nullptr, nullptr, nullptr, gcap.data(), &hd) != GL_OK) return false;
    bool ok = L.tree_read(ctx, hd, GL_PART_COEFFS, gc.data()) == GL_OK && L.tree_read(ctx, hd, GL_PART_LEAVES, gl.data()) == GL_OK &&
              L.tree_read(ctx, hd, GL_PART_DIGESTS, gd.data()) == GL_OK;
    L.tree_free(ctx, hd);
    return ok && gc == rc_ && gl == rl && !std::memcmp(gd.data(), rd.data(), 8 * nd) && gcap == rcap;
}

int child(const char* path, unsigned log_n, unsigned c, unsigned r, const uint64_t* host_in) {
    Lib L;
    if (!L.open(path)) return 1;
    gl_ctx* ctx = nullptr;
    if (L.ctx_create(&ctx, 0) != GL_OK) { std::printf("{\"lib\": \"%s\", \"error\": \"gl_ctx_create failed (no CPU fallback)\"}\n", path); return 2; }
    const bool p1 = parity(L, ctx, 10, 135, 3, 4, true, 2025), p2 = parity(L, ctx, 7, 20, 1, 2, false, 7), p3 = parity(L, ctx, 12, 9, 2, 0, true, 11);
    const uint64_t N = 1ULL << log_n;
    uint64_t* in = (uint64_t*)L.host_alloc(N * c * 8);
    if (!in) { std::printf("{\"lib\": \"%s\", \"error\": \"pinned allocation failed\"}\n", path); return 1; }
    std::memcpy(in, host_in, N * c * 8);
    std::vector<const uint64_t*> cols(c);
    for (unsigned j = 0; j < c; j++) cols[j] = in + (uint64_t)j * N;
    uint64_t cap[64];
    float best[GL_N_STAGES];
    for (auto& b : best) b = 1e30f;
    double best_wall = 1e30;
    for (int it = 0; it < 6; it++) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = L.commit(ctx, cols.data(), c, log_n, r, 4, 0, nullptr, nullptr, nullptr, cap, nullptr);
        auto t1 = std::chrono::steady_clock::now();
        if (rc != GL_OK) { std::printf("{\"lib\": \"%s\", \"error\": \"gl_commit: %s\"}\n", path, L.last_error(ctx)); return 1; }
        if (it < 2) continue;
        float st[GL_N_STAGES];
        L.stage_times(ctx, st, nullptr);
        for (int i = 0; i < GL_N_STAGES; i++) best[i] = std::min(best[i], st[i]);
        best_wall = std::min(best_wall, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
    std::printf("{\"lib\": \"%s\", \"parity_2p10x135\": %s, \"parity_noncanonical\": %s, \"parity_2p12x9\": %s, \"log_n\": %u, \"n_cols\": %u, \"rate_bits\": %u, "
                "\"e2e_ms\": %.3f, \"melem_per_s\": %.1f, \"intt_ms\": %.3f, \"lde_ms\": %.3f, \"leaf_hash_ms\": %.3f, \"tree_ms\": %.3f, \"cap0\": \"%016llx\"}\n",
                path, p1 ? "true" : "false", p2 ? "true" : "false", p3 ? "true" : "false", log_n, c, r, best_wall, (double)N * c / (best_wall * 1e3),
                best[GL_STAGE_INTT], best[GL_STAGE_LDE], best[GL_STAGE_LEAF_HASH], best[GL_STAGE_TREE], (unsigned long long)cap[0]);
    std::fflush(stdout);
    L.ctx_destroy(ctx);
    return (p1 && p2 && p3) ? 0 : 3;
}
}  // namespace

int main(int argc, char** argv) {
    unsigned log_n = 20, c = 135, r = 3;
    std::vector<const char*> libs;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--log-n") && i + 1 < argc) log_n = atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--cols") && i + 1 < argc) c = atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--rate-bits") && i + 1 < argc) r = atoi(argv[++i]);
        else libs.push_back(argv[i]);
    }
    if (libs.empty()) { std::fprintf(stderr, "usage: variant_bench [--log-n L] [--cols C] [--rate-bits R] lib.so ...\n"); return 64; }
    const uint64_t N = 1ULL << log_n;
    std::vector<uint64_t> in(N * c);   // generated once in the parent (which never touches CUDA) and inherited by the children
    {
        std::vector<std::thread> th;
        const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        for (unsigned t = 0; t < T; t++)
            th.emplace_back([&, t] { for (unsigned j = t; j < c; j += T) fill(in.data() + (uint64_t)j * N, N, (uint64_t)j * N, 2025, true); });
        for (auto& x : th) x.join();
    }
    int bad = 0;
    for (const char* lib : libs) {
        std::fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) {
            const int rc = child(lib, log_n, c, r, in.data());
            std::fflush(stdout);
            _exit(rc);
        }
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { bad++; std::printf("{\"lib\": \"%s\", \"exit\": %d}\n", lib, WIFEXITED(st) ? WEXITSTATUS(st) : -1); }
    }
    return bad ? 1 : 0;
}
