// Parity test of the C++ host mirror (include/gl_plonky2.hpp) against the C oracle (oracle/libgl_oracle.so, checker only).
// Reads like upstream's own tests: commit -> prove_openings -> verify every Merkle path with the verifier's rule.
// Built and run by tests/test_zz_cpp_host.py:   host_mirror_test [--expect-no-device] [--golden FILE]
// Exit code 0 = every check passed.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>

#include "gl_plonky2.hpp"

extern "C" {
// oracle/gl_oracle.c
typedef struct {
    uint64_t state[12];
    uint64_t in[8]; int n_in;
    uint64_t out[8]; int n_out;
} glo_challenger;
void glo_poseidon(uint64_t s[12]);
void glo_hash_or_noop(const uint64_t* in, uint64_t len, uint64_t out[4]);
void glo_two_to_one(const uint64_t* l, const uint64_t* r, uint64_t out[4]);
void glo_coset_fft(uint64_t* a, unsigned log_n, uint64_t shift);
int glo_merkle_new(const uint64_t* leaves, uint64_t n_leaves, uint64_t leaf_len, unsigned cap_height, uint64_t* digests, uint64_t* cap);
int glo_commit(const uint64_t* const* cols, uint32_t n_cols, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int is_coeffs,
               uint64_t* out_coeffs, uint64_t* out_leaves, uint64_t* out_digests, uint64_t* out_cap, double* stage_s);
void glo_challenger_init(glo_challenger* c);
void glo_challenger_observe(glo_challenger* c, const uint64_t* e, uint64_t n);
uint64_t glo_challenger_get(glo_challenger* c);
unsigned glo_challenger_sizeof(void);
uint64_t glo_fri_proof_of_work(glo_challenger* c, uint32_t min_leading_zeros);
int glo_fri_committed_trees(const uint64_t* coeffs_in, const uint64_t* values_in, uint64_t len, const uint32_t* arity_bits, uint32_t n_layers,
                            uint32_t rate_bits, uint32_t cap_height, glo_challenger* ch, uint64_t** out_leaves, uint64_t** out_digests,
                            uint64_t** out_caps, uint64_t* out_betas, uint64_t* final_poly);
int glo_openings_add_batch(const uint64_t* const* polys, uint32_t n_polys, uint64_t n, const uint64_t alpha[2], const uint64_t point[2],
                           uint64_t* final_io, uint64_t* quotient_out);
}

using namespace plonky2;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...)                                                  \
    do {                                                                  \
        g_checks++;                                                       \
        if (!(cond)) {                                                    \
            g_fail++;                                                     \
            std::printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond);   \
            std::printf(__VA_ARGS__);                                     \
            std::printf("\n");                                            \
        }                                                                 \
    } while (0)

// tests/oracle_c.py · splitmix_columns
static std::vector<std::vector<F>> splitmix_columns(uint64_t seed, size_t n_cols, size_t n, bool canonical = true) {
    std::vector<std::vector<F>> cols(n_cols, std::vector<F>(n));
    uint64_t idx = 1;
    for (size_t c = 0; c < n_cols; c++)
        for (size_t i = 0; i < n; i++, idx++) {
            uint64_t z = (0x706C6F6E6B7932ULL ^ seed) + idx * 0x9E3779B97F4A7C15ULL;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            z ^= z >> 31;
            if (canonical && z >= ORDER) z -= ORDER;
            cols[c][i] = z;
        }
    return cols;
}

static uint64_t bitrev(uint64_t x, unsigned bits) {
    uint64_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// plonky2 hash/merkle_proofs.rs · verify_merkle_proof_to_cap, with the oracle's hash
static bool verify_path(const std::vector<F>& leaf, uint64_t index, const MerkleProof& proof, const MerkleCap& cap) {
    uint64_t cur[4], nxt[4];
    glo_hash_or_noop(leaf.data(), leaf.size(), cur);
    for (const HashOut& sib : proof.siblings) {
        if (index & 1) glo_two_to_one(sib.elements.data(), cur, nxt);
        else glo_two_to_one(cur, sib.elements.data(), nxt);
        std::memcpy(cur, nxt, 32);
        index >>= 1;
    }
    return index < cap.hashes.size() && std::memcmp(cur, cap.hashes[index].elements.data(), 32) == 0;
}

template <class Fn>
static bool panics_with(const char* needle, Fn&& fn) {
    try { fn(); } catch (const Panic& e) { return std::strstr(e.what(), needle) != nullptr; } catch (...) { return false; }
    return false;
}

struct OracleCommit { std::vector<F> coeffs, leaves, digests, cap; };
static OracleCommit oracle_commit(const std::vector<std::vector<F>>& cols, unsigned log_n, unsigned r, unsigned h, int is_coeffs) {
    const size_t n = size_t(1) << log_n, R = n << r, c = cols.size();
    OracleCommit o;
    o.coeffs.resize(c * n); o.leaves.resize(R * c); o.digests.resize(8 * (R - (size_t(1) << h)) + 4); o.cap.resize(4 << h);
    std::vector<const uint64_t*> p;
    for (auto& v : cols) p.push_back(v.data());
    int rc = glo_commit(p.data(), uint32_t(c), log_n, r, h, is_coeffs, o.coeffs.data(), o.leaves.data(), o.digests.data(), o.cap.data(), nullptr);
    CHECK(rc == 0, "oracle rejected the shape");
    o.digests.resize(8 * (R - (size_t(1) << h)));
    return o;
}

static std::vector<PolynomialValues> as_values(const std::vector<std::vector<F>>& cols) {
    std::vector<PolynomialValues> v(cols.size());
    for (size_t i = 0; i < cols.size(); i++) v[i].values = cols[i];
    return v;
}

static void test_poseidon_and_challenger(Context& ctx) {
    auto st = splitmix_columns(7, 64, 12, false);                 // 64 states incl. non-canonical words
    std::vector<F> flat;
    for (auto& s : st) flat.insert(flat.end(), s.begin(), s.end());
    ctx.poseidon_permute(flat.data(), st.size());
    bool ok = true;
    for (size_t i = 0; i < st.size(); i++) {
        uint64_t s[12];
        for (int k = 0; k < 12; k++) s[k] = st[i][k] % ORDER;
        glo_poseidon(s);
        ok &= std::memcmp(s, flat.data() + 12 * i, 96) == 0;
    }
    CHECK(ok, "PoseidonPermutation::permute differs from the oracle");

    CHECK(glo_challenger_sizeof() == sizeof(glo_challenger), "oracle challenger layout");
    Challenger a(&ctx);
    glo_challenger b;
    glo_challenger_init(&b);
    uint64_t x = 12345;
    ok = true;
    for (int round = 0; round < 12; round++) {
        size_t k = 1 + (x % 12);
        std::vector<F> es(k);
        for (auto& e : es) { x = x * 6364136223846793005ULL + 1442695040888963407ULL; e = x; }   // non-canonical words too
        a.observe_elements(es);
        glo_challenger_observe(&b, es.data(), k);
        for (size_t d = 0; d < (x >> 60) % 3; d++) ok &= a.get_challenge() == glo_challenger_get(&b);
    }
    HashOut h = a.get_hash();
    for (int k = 0; k < 4; k++) ok &= h.elements[k] == glo_challenger_get(&b);
    CHECK(ok, "Challenger transcript differs from the oracle");
}

static void test_merkle(Context& ctx) {
    const unsigned shapes[][3] = {{1, 5, 0}, {2, 1, 1}, {8, 4, 3}, {16, 5, 2}, {64, 8, 3}, {32, 9, 5}, {256, 135, 4}, {1024, 32, 4}, {128, 3, 0}};
    for (auto& s : shapes) {
        const size_t n = s[0], len = s[1], h = s[2];
        auto lv = splitmix_columns(100 + n + len, n, len);
        MerkleTree t = MerkleTree::new_(lv, h, &ctx);
        std::vector<F> flat;
        for (auto& l : lv) flat.insert(flat.end(), l.begin(), l.end());
        std::vector<F> dig(8 * (n - (size_t(1) << h)) + 4), cap(4 << h);
        CHECK(glo_merkle_new(flat.data(), n, len, h, dig.data(), cap.data()) == 0, "oracle merkle_new");
        CHECK(t.cap.flatten() == cap, "cap differs n=%zu len=%zu h=%zu", n, len, h);
        CHECK(t.cap.height() == h && t.n_leaves() == n && t.leaf_len() == len && t.cap_height() == h, "tree info");
        const auto& d = t.digests();
        CHECK(d.size() == 2 * (n - (size_t(1) << h)) && (d.empty() || std::memcmp(d[0].elements.data(), dig.data(), 32 * d.size()) == 0),
              "digests differ n=%zu len=%zu h=%zu", n, len, h);
        CHECK(t.leaves() == flat, "leaves differ");
        bool ok = true;
        for (size_t i = 0; i < n; i += (n > 16 ? n / 8 + 1 : 1)) {
            auto row = t.get(i);
            auto proof = t.prove(i);
            ok &= row == lv[i] && proof.len() == t.depth() && verify_path(row, i, proof, t.cap);
        }
        CHECK(ok, "get/prove/verify n=%zu len=%zu h=%zu", n, len, h);
    }
    auto lv = splitmix_columns(5, 8, 3);
    CHECK(panics_with("cap_height", [&] { MerkleTree::new_(lv, 4, &ctx); }), "cap_height > log2(n) must panic");
    lv.pop_back();
    CHECK(panics_with("", [&] { MerkleTree::new_(lv, 0, &ctx); }), "non power-of-two leaf count must panic");
    auto t = MerkleTree::new_(splitmix_columns(5, 8, 3), 1, &ctx);
    CHECK(panics_with("out of range", [&] { t.get(8); }), "get out of range must panic");
    CHECK(panics_with("out of range", [&] { t.open_batch({0, 8}); }), "open_batch out of range must panic");
    CHECK(t.open_batch({}).empty(), "empty open_batch");
}

static void test_commit(Context& ctx) {
    const unsigned shapes[][4] = {{0, 1, 0, 0}, {0, 3, 2, 1}, {3, 1, 1, 2}, {4, 4, 3, 4}, {5, 7, 1, 0}, {6, 9, 3, 9},
                                  {8, 20, 3, 4}, {10, 135, 3, 4}, {11, 16, 1, 4}, {12, 2, 2, 3}};
    for (auto& s : shapes) {
        const unsigned log_n = s[0], c = s[1], r = s[2], h = s[3];
        const size_t n = size_t(1) << log_n, R = n << r;
        for (int is_coeffs = 0; is_coeffs < 2; is_coeffs++) {
            auto cols = splitmix_columns(500 + 7 * log_n + c, c, n, /*canonical=*/is_coeffs == 0);
            OracleCommit ref = oracle_commit(cols, log_n, r, h, is_coeffs);
            PolynomialBatch pb;
            if (is_coeffs) {
                std::vector<PolynomialCoeffs> p(c);
                for (unsigned j = 0; j < c; j++) p[j].coeffs = cols[j];
                pb = PolynomialBatch::from_coeffs(p, r, false, h, nullptr, nullptr, &ctx);
            } else {
                TimingTree timing;
                pb = PolynomialBatch::from_values(as_values(cols), r, false, h, &timing, nullptr, &ctx);
            }
            CHECK(pb.degree_log == log_n && pb.rate_bits == r && !pb.blinding && pb.num_polys() == c, "batch fields");
            CHECK(pb.merkle_tree.cap.flatten() == ref.cap, "cap differs 2^%u x %u r=%u h=%u coeffs=%d", log_n, c, r, h, is_coeffs);
            const auto& polys = pb.polynomials();
            bool ok = polys.size() == c;
            for (unsigned j = 0; ok && j < c; j++) ok = std::memcmp(polys[j].coeffs.data(), ref.coeffs.data() + j * n, 8 * n) == 0;
            CHECK(ok, "polynomials differ 2^%u x %u coeffs=%d", log_n, c, is_coeffs);
            CHECK(pb.merkle_tree.leaves() == ref.leaves, "leaves differ 2^%u x %u r=%u coeffs=%d", log_n, c, r, is_coeffs);
            const auto& d = pb.merkle_tree.digests();
            CHECK(4 * d.size() == ref.digests.size() && (d.empty() || std::memcmp(d[0].elements.data(), ref.digests.data(), 32 * d.size()) == 0),
                  "digests differ 2^%u x %u r=%u h=%u", log_n, c, r, h);
            // get_lde_values(index, step) = leaves[reverse_bits(index * step)] (what the quotient stage reads)
            const size_t step = size_t(1) << (r ? r - 1 : 0);
            ok = true;
            for (size_t i = 0; i < R / step; i += (R / step > 8 ? R / step / 5 : 1)) {
                auto row = pb.get_lde_values(i, step);
                const size_t src = bitrev(i * step, log_n + r);
                ok &= std::memcmp(row.data(), ref.leaves.data() + src * c, 8 * c) == 0;
            }
            CHECK(ok, "get_lde_values differs");
        }
    }
    auto cols = splitmix_columns(1, 3, 8);
    CHECK(panics_with("zero_knowledge", [&] { PolynomialBatch::from_values(as_values(cols), 1, true, 1, nullptr, nullptr, &ctx); }), "blinding must panic");
    CHECK(panics_with("cap_height", [&] { PolynomialBatch::from_values(as_values(cols), 1, false, 5, nullptr, nullptr, &ctx); }), "cap_height must panic");
    CHECK(panics_with("empty", [&] { PolynomialBatch::from_values({}, 1, false, 0, nullptr, nullptr, &ctx); }), "empty batch must panic");
    cols[1].resize(4);
    CHECK(panics_with("Polynomial degrees inconsistent", [&] { PolynomialBatch::from_values(as_values(cols), 1, false, 1, nullptr, nullptr, &ctx); }),
          "ragged columns must panic");
}

struct OracleFri { std::vector<std::vector<F>> leaves, digests, caps; std::vector<F> final_poly; };
static OracleFri oracle_fri(const std::vector<F>& coeffs, const std::vector<F>& values, size_t len, const std::vector<size_t>& arity_bits, unsigned r,
                            unsigned h, glo_challenger* ch) {
    OracleFri o;
    std::vector<uint32_t> ab(arity_bits.begin(), arity_bits.end());
    size_t cur = len, total = 0;
    std::vector<uint64_t*> pl, pd, pc;
    for (size_t a : arity_bits) {
        size_t nl = cur >> a;
        o.leaves.emplace_back(2 * cur);
        o.digests.emplace_back(8 * (nl - (size_t(1) << h)) + 4);
        o.caps.emplace_back(4 << h);
        cur = nl; total += a;
    }
    for (size_t l = 0; l < arity_bits.size(); l++) { pl.push_back(o.leaves[l].data()); pd.push_back(o.digests[l].data()); pc.push_back(o.caps[l].data()); }
    std::vector<F> betas(2 * arity_bits.size() + 2);
    o.final_poly.resize(2 * ((len >> total) >> r) + 2);
    int rc = glo_fri_committed_trees(coeffs.data(), values.data(), len, ab.data(), uint32_t(ab.size()), r, h, ch, pl.data(), pd.data(), pc.data(),
                                     betas.data(), o.final_poly.data());
    CHECK(rc == 0, "oracle fri_committed_trees");
    o.final_poly.resize(2 * ((len >> total) >> r));
    return o;
}

static void test_fri_proof_host_arrays(Context& ctx) {
    struct Cfg { unsigned log_len, r, h; std::vector<size_t> arity; uint32_t pow_bits; size_t rounds; };
    const Cfg cfgs[] = {{8, 3, 1, {4}, 5, 4}, {9, 1, 2, {2, 3}, 0, 3}, {12, 3, 4, {4, 4}, 8, 28}, {15, 3, 4, {4, 4}, 10, 28}};
    for (const Cfg& cfg : cfgs) {
        const size_t n = size_t(1) << cfg.log_len, low = n >> cfg.r;
        auto raw = splitmix_columns(cfg.log_len, n, 2);
        PolynomialCoeffsExt co; PolynomialValuesExt va;
        co.coeffs.resize(n); va.values.resize(n);
        std::vector<F> c0(n, 0), c1(n, 0);
        for (size_t i = 0; i < low; i++) { co.coeffs[i] = Ext{raw[i][0], raw[i][1]}; c0[i] = raw[i][0]; c1[i] = raw[i][1]; }
        for (size_t i = low; i < n; i++) co.coeffs[i] = Ext{0, 0};
        glo_coset_fft(c0.data(), cfg.log_len, 7);
        glo_coset_fft(c1.data(), cfg.log_len, 7);
        for (size_t i = 0; i < n; i++) va.values[i] = Ext{c0[i], c1[i]};
        std::vector<F> cof(2 * n), vaf(2 * n);
        std::memcpy(cof.data(), co.coeffs.data(), 16 * n);
        std::memcpy(vaf.data(), va.values.data(), 16 * n);

        glo_challenger ref;
        glo_challenger_init(&ref);
        Challenger ch(&ctx);
        const uint64_t prefix[3] = {1, 2, 3};
        glo_challenger_observe(&ref, prefix, 3);
        ch.observe_elements(prefix, 3);
        OracleFri o = oracle_fri(cof, vaf, n, cfg.arity, cfg.r, cfg.h, &ref);
        const uint64_t w_ref = glo_fri_proof_of_work(&ref, cfg.pow_bits);

        FriParams p;
        p.config.rate_bits = cfg.r; p.config.cap_height = cfg.h; p.config.proof_of_work_bits = cfg.pow_bits; p.config.num_query_rounds = cfg.rounds;
        p.degree_bits = cfg.log_len - cfg.r;
        p.reduction_arity_bits = cfg.arity;
        // one initial oracle so that initial_trees_proof is exercised: any tree over n leaves
        auto lv = splitmix_columns(77, n, 6);
        MerkleTree initial = MerkleTree::new_(lv, cfg.h, &ctx);
        FriProof proof = fri_proof({&initial}, co, va, ch, p, nullptr, &ctx);

        CHECK(proof.commit_phase_merkle_caps.size() == cfg.arity.size(), "one cap per layer");
        for (size_t l = 0; l < cfg.arity.size(); l++)
            CHECK(proof.commit_phase_merkle_caps[l].flatten() == o.caps[l], "commit-phase cap %zu differs (2^%u)", l, cfg.log_len);
        CHECK(proof.final_poly.len() * 2 == o.final_poly.size() &&
                  std::memcmp(proof.final_poly.coeffs.data(), o.final_poly.data(), 8 * o.final_poly.size()) == 0, "final_poly differs");
        CHECK(proof.pow_witness == w_ref, "pow_witness %" PRIu64 " vs oracle %" PRIu64, proof.pow_witness, w_ref);
        CHECK(proof.query_round_proofs.size() == cfg.rounds, "num_query_rounds");
        bool ok = true;
        for (const FriQueryRound& rnd : proof.query_round_proofs) {
            uint64_t x = glo_challenger_get(&ref) % n;                       // the verifier's replay of the transcript
            const auto& ip = rnd.initial_trees_proof.evals_proofs;
            ok &= ip.size() == 1 && ip[0].first == lv[x] && verify_path(ip[0].first, x, ip[0].second, initial.cap);
            ok &= rnd.steps.size() == cfg.arity.size();
            for (size_t l = 0; ok && l < cfg.arity.size(); l++) {
                const size_t ab = cfg.arity[l], arity = size_t(1) << ab, leaf_index = x >> ab;
                std::vector<F> leaf(o.leaves[l].begin() + leaf_index * 2 * arity, o.leaves[l].begin() + (leaf_index + 1) * 2 * arity);
                std::vector<Ext> expect;
                for (size_t k = 0; k < arity; k++) expect.push_back(Ext{leaf[2 * k], leaf[2 * k + 1]});   // validate_shape: evals.len() == arity
                ok &= rnd.steps[l].evals == expect;
                MerkleCap cap;
                cap.hashes.resize(size_t(1) << cfg.h);
                std::memcpy(cap.hashes[0].elements.data(), o.caps[l].data(), 32 << cfg.h);
                ok &= verify_path(leaf, leaf_index, rnd.steps[l].merkle_proof, cap);
                x = leaf_index;
            }
        }
        CHECK(ok, "query rounds (2^%u)", cfg.log_len);
        CHECK(ch.get_challenge() == glo_challenger_get(&ref), "transcripts diverged after fri_proof");
    }
}

using Golden = std::map<std::string, std::vector<uint64_t>>;
static Golden read_golden(const char* path) {
    Golden g;
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        uint64_t v;
        while (ss >> v) g[key].push_back(v);
    }
    return g;
}

// commits -> PolynomialBatch::prove_openings on the device vs the oracle on the same transcript (plonky2-shaped instance:
// every polynomial at zeta, the first two of the last oracle also at g*zeta), then every path with the verifier's rule
static void test_prove_openings(Context& ctx, const Golden* golden) {
    struct Cfg { unsigned log_n; std::vector<size_t> widths; unsigned r, h; std::vector<size_t> arity; uint32_t pow_bits; size_t rounds;
                 std::vector<uint64_t> seeds, prefix; Ext zeta, gz; };
    std::vector<Cfg> cfgs = {
        {3, {3, 2}, 1, 1, {}, 4, 3, {900, 917}, {5, 6, 7, 8, 9}, {111, 222}, {333, 444}},
        {6, {9, 5, 2, 2}, 3, 2, {2}, 6, 5, {1, 2, 3, 4}, {5, 6, 7, 8, 9}, {0xFFFFFFFF00000000ULL, 17}, {3, 0}},
        {10, {135, 20}, 3, 4, {4, 4}, 8, 28, {10, 11}, {5}, {123456789, 987654321}, {5, 0}},
        {12, {86, 135, 20, 16}, 3, 4, {4, 4}, 10, 28, {20, 21, 22, 23}, {1, 2, 3, 4, 5, 6, 7, 8, 9}, {42, 43}, {44, 45}},
    };
    if (golden) {   // tests/golden/openings_fri.json, flattened by tests/test_zz_cpp_host.py
        const Golden& g = *golden;
        Cfg c;
        c.log_n = unsigned(g.at("log_n")[0]); c.r = unsigned(g.at("rate_bits")[0]); c.h = unsigned(g.at("cap_height")[0]);
        for (auto v : g.at("widths")) c.widths.push_back(v);
        for (auto v : g.at("arity_bits")) c.arity.push_back(v);
        c.pow_bits = uint32_t(g.at("pow_bits")[0]); c.rounds = g.at("query_indices").size();
        c.seeds = g.at("col_seeds"); c.prefix = g.at("transcript_prefix");
        c.zeta = Ext{g.at("point0")[0], g.at("point0")[1]}; c.gz = Ext{g.at("point1")[0], g.at("point1")[1]};
        cfgs.push_back(c);
    }
    for (size_t ci = 0; ci < cfgs.size(); ci++) {
        const Cfg& cfg = cfgs[ci];
        const bool is_golden = golden && ci + 1 == cfgs.size();
        const size_t n = size_t(1) << cfg.log_n, R = n << cfg.r;
        std::vector<PolynomialBatch> batches;
        for (size_t k = 0; k < cfg.widths.size(); k++)
            batches.push_back(PolynomialBatch::from_values(as_values(splitmix_columns(cfg.seeds[k], cfg.widths[k], n)), cfg.r, false, cfg.h, nullptr, nullptr, &ctx));
        std::vector<const PolynomialBatch*> oracles;
        for (auto& b : batches) oracles.push_back(&b);
        FriInstanceInfo instance;
        FriBatchInfo all{cfg.zeta, {}}, some{cfg.gz, {}};
        for (size_t k = 0; k < cfg.widths.size(); k++)
            for (size_t i = 0; i < cfg.widths[k]; i++) all.polynomials.push_back({k, i});
        for (size_t i = 0; i < std::min<size_t>(2, cfg.widths.back()); i++) some.polynomials.push_back({cfg.widths.size() - 1, i});
        instance.batches = {all, some};
        for (auto w : cfg.widths) instance.oracles.push_back({w, false});

        glo_challenger ref;
        glo_challenger_init(&ref);
        Challenger ch(&ctx);
        glo_challenger_observe(&ref, cfg.prefix.data(), cfg.prefix.size());
        ch.observe_elements(cfg.prefix);

        // oracle side: alpha, final_poly, lde, coset_fft, commit phase, PoW
        uint64_t alpha[2];
        alpha[0] = glo_challenger_get(&ref); alpha[1] = glo_challenger_get(&ref);
        std::vector<F> final_poly(2 * n, 0);
        for (const FriBatchInfo& b : instance.batches) {
            std::vector<const uint64_t*> polys;
            for (auto& fpi : b.polynomials) polys.push_back(oracles[fpi.oracle_index]->polynomials()[fpi.polynomial_index].coeffs.data());
            CHECK(glo_openings_add_batch(polys.data(), uint32_t(polys.size()), n, alpha, b.point.data(), final_poly.data(), nullptr) == 0, "oracle add_batch");
        }
        std::vector<F> c0(R, 0), c1(R, 0), lde(2 * R, 0), vals(2 * R);
        for (size_t i = 0; i < n; i++) { c0[i] = final_poly[2 * i]; c1[i] = final_poly[2 * i + 1]; lde[2 * i] = c0[i]; lde[2 * i + 1] = c1[i]; }
        glo_coset_fft(c0.data(), cfg.log_n + cfg.r, 7);
        glo_coset_fft(c1.data(), cfg.log_n + cfg.r, 7);
        for (size_t i = 0; i < R; i++) { vals[2 * i] = c0[i]; vals[2 * i + 1] = c1[i]; }
        OracleFri o = oracle_fri(lde, vals, R, cfg.arity, cfg.r, cfg.h, &ref);
        const uint64_t w_ref = glo_fri_proof_of_work(&ref, cfg.pow_bits);

        FriParams p;
        p.config.rate_bits = cfg.r; p.config.cap_height = cfg.h; p.config.proof_of_work_bits = cfg.pow_bits; p.config.num_query_rounds = cfg.rounds;
        p.degree_bits = cfg.log_n;
        p.reduction_arity_bits = cfg.arity;
        FriProof proof = PolynomialBatch::prove_openings(instance, oracles, ch, p, nullptr, &ctx);

        for (size_t l = 0; l < cfg.arity.size(); l++)
            CHECK(proof.commit_phase_merkle_caps[l].flatten() == o.caps[l], "cfg %zu: commit-phase cap %zu differs", ci, l);
        CHECK(proof.final_poly.len() * 2 == o.final_poly.size() &&
                  std::memcmp(proof.final_poly.coeffs.data(), o.final_poly.data(), 8 * o.final_poly.size()) == 0, "cfg %zu: final_poly differs", ci);
        CHECK(proof.pow_witness == w_ref, "cfg %zu: pow_witness", ci);
        CHECK(proof.query_round_proofs.size() == cfg.rounds, "cfg %zu: rounds", ci);
        bool ok = true;
        std::vector<uint64_t> xs;
        for (const FriQueryRound& rnd : proof.query_round_proofs) {
            uint64_t x = glo_challenger_get(&ref) % R;
            xs.push_back(x);
            const auto& ip = rnd.initial_trees_proof.evals_proofs;
            ok &= ip.size() == batches.size();
            for (size_t k = 0; ok && k < batches.size(); k++) {
                const auto& lv = batches[k].merkle_tree.leaves();
                const size_t w = cfg.widths[k];
                ok &= ip[k].first == std::vector<F>(lv.begin() + x * w, lv.begin() + (x + 1) * w);
                ok &= verify_path(ip[k].first, x, ip[k].second, batches[k].merkle_tree.cap);
            }
            for (size_t l = 0; ok && l < cfg.arity.size(); l++) {
                const size_t ab = cfg.arity[l], arity = size_t(1) << ab, leaf_index = x >> ab;
                std::vector<F> leaf(o.leaves[l].begin() + leaf_index * 2 * arity, o.leaves[l].begin() + (leaf_index + 1) * 2 * arity);
                std::vector<Ext> expect;
                for (size_t k = 0; k < arity; k++) expect.push_back(Ext{leaf[2 * k], leaf[2 * k + 1]});
                ok &= rnd.steps[l].evals == expect && verify_path(leaf, leaf_index, rnd.steps[l].merkle_proof, proof.commit_phase_merkle_caps[l]);
                x = leaf_index;
            }
        }
        CHECK(ok, "cfg %zu: query rounds", ci);
        CHECK(ch.get_challenge() == glo_challenger_get(&ref), "cfg %zu: transcripts diverged after prove_openings", ci);
        if (is_golden) {
            const Golden& g = *golden;
            CHECK(Ext({alpha[0], alpha[1]}) == Ext({g.at("alpha")[0], g.at("alpha")[1]}), "golden alpha");
            std::vector<F> caps;
            for (auto& c : proof.commit_phase_merkle_caps) { auto f = c.flatten(); caps.insert(caps.end(), f.begin(), f.end()); }
            CHECK(caps == g.at("commit_phase_caps"), "golden commit_phase_caps");
            std::vector<F> fp;
            for (auto& e : proof.final_poly.coeffs) { fp.push_back(e[0]); fp.push_back(e[1]); }
            CHECK(fp == g.at("fri_final_poly"), "golden fri_final_poly");
            CHECK(proof.pow_witness == g.at("pow_witness")[0], "golden pow_witness");
            CHECK(xs == g.at("query_indices"), "golden query_indices");
        }
    }
    // misuse: oracles of different degrees, polynomial index out of range
    auto a = PolynomialBatch::from_values(as_values(splitmix_columns(1, 3, 8)), 1, false, 1, nullptr, nullptr, &ctx);
    auto b = PolynomialBatch::from_values(as_values(splitmix_columns(2, 3, 16)), 1, false, 1, nullptr, nullptr, &ctx);
    FriParams p;
    p.config.rate_bits = 1; p.config.cap_height = 1; p.config.proof_of_work_bits = 0; p.config.num_query_rounds = 1;
    Challenger ch(&ctx);
    FriInstanceInfo bad1; bad1.batches = {{Ext{1, 2}, {{0, 0}, {1, 0}}}};
    CHECK(panics_with("Polynomial degrees inconsistent", [&] { PolynomialBatch::prove_openings(bad1, {&a, &b}, ch, p, nullptr, &ctx); }), "degree mismatch must panic");
    FriInstanceInfo bad2; bad2.batches = {{Ext{1, 2}, {{0, 3}}}};
    CHECK(panics_with("out of range", [&] { PolynomialBatch::prove_openings(bad2, {&a}, ch, p, nullptr, &ctx); }), "polynomial index must panic");
}

// `cargo test` proves from several threads at once: every thread gets its own Context (Context::thread_default(), as the
// Rust shim's thread_local CTX) and commits / opens concurrently; results must not depend on what the other threads do.
static void test_concurrent_contexts() {
    constexpr int T = 4;
    int bad[T] = {0, 0, 0, 0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([t, &bad] {
            try {
                for (int it = 0; it < 3; it++) {
                    const unsigned log_n = 7 + t, c = 20 + 5 * t + it, r = 1 + (t & 1) * 2, h = 2;
                    const size_t n = size_t(1) << log_n, R = n << r;
                    auto cols = splitmix_columns(4000 + 10 * t + it, c, n);
                    std::vector<const uint64_t*> p;
                    for (auto& v : cols) p.push_back(v.data());
                    std::vector<F> leaves(R * c), digests(8 * (R - (size_t(1) << h))), cap(4 << h);
                    if (glo_commit(p.data(), c, log_n, r, h, 0, nullptr, leaves.data(), digests.data(), cap.data(), nullptr)) bad[t]++;
                    PolynomialBatch pb = PolynomialBatch::from_values(as_values(cols), r, false, h);   // Context::thread_default()
                    if (pb.merkle_tree.cap.flatten() != cap) bad[t]++;
                    const size_t x = (R / 3) ^ size_t(t);
                    auto opened = pb.merkle_tree.open_batch({x, R - 1});
                    if (opened[0].first != std::vector<F>(leaves.begin() + x * c, leaves.begin() + (x + 1) * c)) bad[t]++;
                    if (!verify_path(opened[0].first, x, opened[0].second, pb.merkle_tree.cap)) bad[t]++;
                    if (!verify_path(opened[1].first, R - 1, opened[1].second, pb.merkle_tree.cap)) bad[t]++;
                    Challenger ch;                                                              // same thread-local context
                    ch.observe_cap(pb.merkle_tree.cap);
                    FriConfig cfg; cfg.proof_of_work_bits = 6;
                    const F w = fri_proof_of_work(ch, cfg);
                    glo_challenger ref;
                    glo_challenger_init(&ref);
                    glo_challenger_observe(&ref, cap.data(), cap.size());
                    if (w != glo_fri_proof_of_work(&ref, 6)) bad[t]++;
                }
            } catch (const std::exception& e) {
                std::printf("thread %d: %s\n", t, e.what());
                bad[t] += 100;
            }
        });
    for (auto& x : th) x.join();
    for (int t = 0; t < T; t++) CHECK(bad[t] == 0, "thread %d: %d mismatches", t, bad[t]);
}

// FriConfig::fri_params: the wrapper circuit (degree_bits 16, standard_recursion_config) folds 4, 4, 4 down to 16 coefficients
// (SURVEY.md A.7); host arithmetic only
static void test_fri_params() {
    FriConfig cfg;
    FriParams p = cfg.fri_params(16, false);
    CHECK(p.reduction_arity_bits == std::vector<size_t>({4, 4, 4}) && p.final_poly_len() == 16 && p.lde_bits() == 19, "standard config, degree_bits 16");
    CHECK(cfg.fri_params(5, false).reduction_arity_bits.empty(), "degree_bits <= final_poly_bits: no reduction");
    CHECK(cfg.fri_params(12, false).reduction_arity_bits == std::vector<size_t>({4, 4}), "degree_bits 12");
    FriConfig tall = cfg;
    tall.cap_height = 14;                              // a layer must keep at least 2^cap_height leaves
    CHECK(tall.fri_params(16, false).reduction_arity_bits == std::vector<size_t>({4}), "cap-limited reduction");
    FriConfig fixed = cfg;
    fixed.reduction_strategy.kind = FriReductionStrategy::Fixed;
    fixed.reduction_strategy.fixed = {3, 2};
    CHECK(fixed.fri_params(10, false).reduction_arity_bits == std::vector<size_t>({3, 2}), "Fixed strategy");
}

// from_values_multi: n contexts (all on device 0 here), the batch as shard trees under one cap
static void test_commit_multi() {
    const unsigned cfgs[][5] = {{2, 8, 20, 1, 1}, {4, 10, 135, 3, 4}, {8, 7, 9, 3, 3}};
    for (auto& cfg : cfgs) {
        const unsigned world = cfg[0], log_n = cfg[1], c = cfg[2], r = cfg[3], h = cfg[4];
        const size_t n = size_t(1) << log_n, R = n << r;
        std::vector<std::unique_ptr<Context>> own;
        std::vector<Context*> ctxs;
        for (unsigned k = 0; k < world; k++) { own.push_back(std::make_unique<Context>(0)); ctxs.push_back(own.back().get()); }
        auto cols = splitmix_columns(800 + world, c, n);
        OracleCommit ref = oracle_commit(cols, log_n, r, h, 0);
        auto sharded = PolynomialBatch::from_values_multi(ctxs, as_values(cols), r, false, h);
        CHECK(sharded.cap.flatten() == ref.cap, "multi: cap differs (world %u)", world);
        CHECK(sharded.shards.size() == world && sharded.rows_per_shard() == R / world, "multi: shard shape");
        bool ok = true;
        for (size_t i = 0; i < R; i += R / 7 + 1) {
            auto row = sharded.get(i);
            ok &= std::memcmp(row.data(), ref.leaves.data() + i * c, 8 * c) == 0 && verify_path(row, i, sharded.prove(i), sharded.cap);
        }
        CHECK(ok, "multi: rows / paths (world %u)", world);
        size_t off = 0;
        ok = true;
        for (auto& t : sharded.shards) {
            const auto& d = t.digests();
            ok &= d.empty() || std::memcmp(d[0].elements.data(), ref.digests.data() + off, 32 * d.size()) == 0;
            off += 4 * d.size();
        }
        CHECK(ok && off == ref.digests.size(), "multi: digests (world %u)", world);
        CHECK(panics_with("twice", [&] { PolynomialBatch::from_values_multi({ctxs[0], ctxs[0]}, as_values(cols), r, false, h); }), "same context twice must panic");
        CHECK(panics_with("cap_height too small", [&] { PolynomialBatch::from_values_multi(ctxs, as_values(cols), r, false, 0); }), "cap_height < log2(world) must panic");
    }
}

int main(int argc, char** argv) {
    bool expect_no_device = false;
    const char* golden_path = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--expect-no-device")) expect_no_device = true;
        else if (!std::strcmp(argv[i], "--golden") && i + 1 < argc) golden_path = argv[++i];
    }
    if (expect_no_device) {   // the product path must fail loudly, not fall back to a CPU path
        try {
            Context c(0);
        } catch (const GlError& e) {
            std::printf("ok: no device -> GlError(%d): %s\n", e.code, e.what());
            return e.code == GL_ERR_CUDA ? 0 : 1;
        }
        std::printf("FAIL: Context(0) succeeded although no device was expected\n");
        return 1;
    }
    try {
        Context ctx(0);
        Golden golden;
        if (golden_path) golden = read_golden(golden_path);
        test_poseidon_and_challenger(ctx);
        std::printf("poseidon/challenger: %d checks, %d failed\n", g_checks, g_fail);
        test_merkle(ctx);
        std::printf("merkle: %d checks, %d failed\n", g_checks, g_fail);
        test_commit(ctx);
        std::printf("commit: %d checks, %d failed\n", g_checks, g_fail);
        test_fri_proof_host_arrays(ctx);
        std::printf("fri_proof: %d checks, %d failed\n", g_checks, g_fail);
        test_prove_openings(ctx, golden_path ? &golden : nullptr);
        std::printf("prove_openings: %d checks, %d failed\n", g_checks, g_fail);
        test_concurrent_contexts();
        std::printf("concurrent contexts: %d checks, %d failed\n", g_checks, g_fail);
        test_commit_multi();
        std::printf("commit_multi: %d checks, %d failed\n", g_checks, g_fail);
        test_fri_params();
        std::printf("fri_params: %d checks, %d failed\n", g_checks, g_fail);
    } catch (const std::exception& e) {
        std::printf("FAIL: unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "ALL OK", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
