// TEST DOUBLE of the C ABI in include/gl_commit.h — test infrastructure, never shipped, never loadable by the product.
//
// It implements the subset of gl_* entry points that include/gl_plonky2.hpp calls, on top of the C oracle
// (oracle/libgl_oracle.so), and is linked STATICALLY into tests/cpp/build/host_mirror_test_double only.  Its one purpose:
// let the CPU test-suite (-m "not gpu", no CUDA device in the build container) exercise the HOST LOGIC of the C++ mirror —
// handle lifetimes, transcript order, query-round assembly, panic mapping — before the same test source runs against the
// real libgl_commit on a B200.  Nothing here is a fallback: libgl_commit.so has no reference to this file, the Python
// package cannot load it, and the GPU tests link the real library.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "gl_commit.h"

extern "C" {
typedef struct {
    uint64_t state[12];
    uint64_t in[8]; int n_in;
    uint64_t out[8]; int n_out;
} glo_challenger;
uint64_t glo_mul(uint64_t a, uint64_t b);
void glo_poseidon(uint64_t s[12]);
void glo_coset_fft(uint64_t* a, unsigned log_n, uint64_t shift);
int glo_merkle_new(const uint64_t* leaves, uint64_t n_leaves, uint64_t leaf_len, unsigned cap_height, uint64_t* digests, uint64_t* cap);
int glo_commit(const uint64_t* const* cols, uint32_t n_cols, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int is_coeffs,
               uint64_t* out_coeffs, uint64_t* out_leaves, uint64_t* out_digests, uint64_t* out_cap, double* stage_s);
void glo_challenger_init(glo_challenger* c);
uint64_t glo_fri_proof_of_work(glo_challenger* c, uint32_t min_leading_zeros);
int glo_openings_add_batch(const uint64_t* const* polys, uint32_t n_polys, uint64_t n, const uint64_t alpha[2], const uint64_t point[2],
                           uint64_t* final_io, uint64_t* quotient_out);
}

namespace {
constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
uint64_t add(uint64_t a, uint64_t b) { unsigned __int128 s = (unsigned __int128)a + b; return uint64_t(s % P); }
void ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t o[2]) {
    uint64_t o0 = add(glo_mul(a[0], b[0]), glo_mul(7, glo_mul(a[1], b[1])));
    uint64_t o1 = add(glo_mul(a[0], b[1]), glo_mul(a[1], b[0]));
    o[0] = o0; o[1] = o1;
}
unsigned log2u(uint64_t n) { unsigned l = 0; while ((1ULL << l) < n) l++; return l; }
uint64_t bitrev(uint64_t x, unsigned bits) { uint64_t r = 0; for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i); return r; }

struct Tree {
    uint64_t n_leaves = 0; uint32_t leaf_len = 0, cap_height = 0, degree_log = 0, rate_bits = 0;
    std::vector<uint64_t> coeffs, leaves, digests, cap;
};
struct Fri { std::vector<uint64_t> coeffs, values; uint64_t len = 0, shift = 7; uint32_t rate_bits = 0, cap_height = 0, arity_bits = 0; };
struct Openings { uint32_t log_n = 0; std::vector<uint64_t> final_poly; };
}  // namespace

struct gl_ctx {
    std::string err;
    uint64_t next = 1;
    std::map<uint64_t, std::unique_ptr<Tree>> trees;
    std::map<uint64_t, std::unique_ptr<Fri>> fris;
    std::map<uint64_t, std::unique_ptr<Openings>> openings;
    int fail(int code, const char* fmt, ...) {
        char buf[256];
        va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        return code;
    }
};

extern "C" {
int gl_abi_version(void) { return GL_ABI_VERSION; }
const char* gl_strerror(int code) {
    switch (code) {
        case GL_OK: return "ok";
        case GL_ERR_INVALID: return "invalid argument";
        case GL_ERR_CUDA: return "CUDA error";
        case GL_ERR_OOM: return "out of memory";
        case GL_ERR_HANDLE: return "unknown handle";
        default: return "unsupported";
    }
}
int gl_ctx_create(gl_ctx** out, int) { *out = new gl_ctx; return GL_OK; }
void gl_ctx_destroy(gl_ctx* c) { delete c; }
const char* gl_ctx_last_error(gl_ctx* c) { return c ? c->err.c_str() : ""; }

int gl_poseidon_permute(gl_ctx*, uint64_t* states, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        for (int k = 0; k < 12; k++) states[12 * i + k] %= P;
        glo_poseidon(states + 12 * i);
    }
    return GL_OK;
}

int gl_poseidon_absorb(gl_ctx*, uint64_t state[12], const uint64_t* groups, uint32_t n_groups) {
    for (uint32_t g = 0; g < n_groups; g++) {
        for (int k = 0; k < 8; k++) state[k] = groups[8 * g + k] % P;
        for (int k = 8; k < 12; k++) state[k] %= P;
        glo_poseidon(state);
    }
    return GL_OK;
}

int gl_commit(gl_ctx* c, const uint64_t* const* cols, uint32_t n_cols, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, int input_is_coeffs,
              uint64_t* out_coeffs, uint64_t* out_leaves, uint64_t* out_digests, uint64_t* out_cap, gl_handle* out_batch) {
    if (!n_cols) return c->fail(GL_ERR_INVALID, "empty polynomial batch");
    if (cap_height > log_n + rate_bits) return c->fail(GL_ERR_INVALID, "cap_height should be at most log2(leaves.len())");
    auto t = std::make_unique<Tree>();
    const uint64_t n = 1ULL << log_n, R = n << rate_bits;
    t->n_leaves = R; t->leaf_len = n_cols; t->cap_height = cap_height; t->degree_log = log_n; t->rate_bits = rate_bits;
    t->coeffs.resize(n * n_cols); t->leaves.resize(R * n_cols); t->digests.resize(8 * (R - (1ULL << cap_height)) + 4); t->cap.resize(4ULL << cap_height);
    if (glo_commit(cols, n_cols, log_n, rate_bits, cap_height, input_is_coeffs, t->coeffs.data(), t->leaves.data(), t->digests.data(), t->cap.data(), nullptr))
        return c->fail(GL_ERR_INVALID, "oracle rejected the shape");
    t->digests.resize(8 * (R - (1ULL << cap_height)));
    if (out_coeffs) std::memcpy(out_coeffs, t->coeffs.data(), 8 * t->coeffs.size());
    if (out_leaves) std::memcpy(out_leaves, t->leaves.data(), 8 * t->leaves.size());
    if (out_digests) std::memcpy(out_digests, t->digests.data(), 8 * t->digests.size());
    std::memcpy(out_cap, t->cap.data(), 8 * t->cap.size());
    if (out_batch) { *out_batch = c->next; c->trees[c->next++] = std::move(t); }
    return GL_OK;
}

// the sharded commit restated on the oracle: one commit, then every context gets its leaf range and its slice of the digests
int gl_commit_multi(gl_ctx* const* ctxs, uint32_t n_ctx, const uint64_t* const* cols, uint32_t n_cols, uint32_t log_n, uint32_t rate_bits,
                    uint32_t cap_height, int input_is_coeffs, uint64_t* out_cap, gl_handle* out_trees) {
    if (!ctxs || !n_ctx || !ctxs[0]) return GL_ERR_INVALID;
    gl_ctx* c = ctxs[0];
    if (n_ctx & (n_ctx - 1)) return c->fail(GL_ERR_INVALID, "n_ctx must be a power of two <= 16");
    for (uint32_t g = 0; g < n_ctx; g++)
        for (uint32_t q = 0; q < g; q++)
            if (ctxs[q] == ctxs[g]) return c->fail(GL_ERR_INVALID, "the same context appears twice");
    if (cap_height > log_n + rate_bits) return c->fail(GL_ERR_INVALID, "cap_height should be at most log2(leaves.len())");
    if ((1ULL << cap_height) < n_ctx) return c->fail(GL_ERR_INVALID, "cap_height too small: every context must own at least one whole cap subtree");
    const uint64_t n = 1ULL << log_n, R = n << rate_bits, rows = R / n_ctx;
    const uint32_t local_h = cap_height - log2u(n_ctx);
    std::vector<uint64_t> leaves(R * n_cols), digests(8 * (R - (1ULL << cap_height)) + 4);
    if (glo_commit(cols, n_cols, log_n, rate_bits, cap_height, input_is_coeffs, nullptr, leaves.data(), digests.data(), out_cap, nullptr))
        return c->fail(GL_ERR_INVALID, "oracle rejected the shape");
    const uint64_t dig_per = 8 * (rows - (1ULL << local_h));
    for (uint32_t g = 0; g < n_ctx; g++) {
        auto t = std::make_unique<Tree>();
        t->n_leaves = rows; t->leaf_len = n_cols; t->cap_height = local_h;
        t->leaves.assign(leaves.begin() + g * rows * n_cols, leaves.begin() + (g + 1) * rows * n_cols);
        t->digests.assign(digests.begin() + g * dig_per, digests.begin() + (g + 1) * dig_per);
        t->cap.assign(out_cap + g * (4ULL << local_h), out_cap + (g + 1) * (4ULL << local_h));
        out_trees[g] = ctxs[g]->next;
        ctxs[g]->trees[ctxs[g]->next++] = std::move(t);
    }
    return GL_OK;
}

static int new_tree(gl_ctx* c, const uint64_t* leaves, uint64_t n_leaves, uint32_t leaf_len, uint32_t cap_height, uint64_t* out_digests, uint64_t* out_cap,
                    gl_handle* out_tree) {
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return c->fail(GL_ERR_INVALID, "n_leaves must be a power of two");
    if ((1ULL << cap_height) > n_leaves) return c->fail(GL_ERR_INVALID, "cap_height should be at most log2(leaves.len())");
    auto t = std::make_unique<Tree>();
    t->n_leaves = n_leaves; t->leaf_len = leaf_len; t->cap_height = cap_height;
    t->leaves.assign(leaves, leaves + n_leaves * leaf_len);
    t->digests.resize(8 * (n_leaves - (1ULL << cap_height)) + 4); t->cap.resize(4ULL << cap_height);
    glo_merkle_new(t->leaves.data(), n_leaves, leaf_len, cap_height, t->digests.data(), t->cap.data());
    t->digests.resize(8 * (n_leaves - (1ULL << cap_height)));
    if (out_digests) std::memcpy(out_digests, t->digests.data(), 8 * t->digests.size());
    if (out_cap) std::memcpy(out_cap, t->cap.data(), 8 * t->cap.size());
    if (out_tree) { *out_tree = c->next; c->trees[c->next++] = std::move(t); }
    return GL_OK;
}
int gl_merkle_new(gl_ctx* c, const uint64_t* leaves, uint64_t n_leaves, uint32_t leaf_len, uint32_t cap_height, uint64_t* out_digests, uint64_t* out_cap,
                  gl_handle* out_tree) {
    return new_tree(c, leaves, n_leaves, leaf_len, cap_height, out_digests, out_cap, out_tree);
}

#define TREE(t, h)                                                        \
    auto it_ = c->trees.find(h);                                          \
    if (it_ == c->trees.end()) return c->fail(GL_ERR_HANDLE, "unknown handle"); \
    Tree* t = it_->second.get()

int gl_tree_info(gl_ctx* c, gl_handle h, gl_tree_info_t* out) {
    TREE(t, h);
    *out = gl_tree_info_t{t->n_leaves, t->leaf_len, t->cap_height, t->degree_log, t->rate_bits, uint32_t(!t->coeffs.empty()), t->leaf_len};
    return GL_OK;
}
int gl_tree_get(gl_ctx* c, gl_handle h, uint64_t i, uint64_t* out_row) {
    TREE(t, h);
    if (i >= t->n_leaves) return c->fail(GL_ERR_INVALID, "leaf index out of range");
    std::memcpy(out_row, t->leaves.data() + i * t->leaf_len, 8 * t->leaf_len);
    return GL_OK;
}
// node j of layer i (0 = leaf digests) of a subtree sits at 2*(((j>>1) << (i+1)) + 2^i - 1) + (j&1)
static uint64_t digest_index(unsigned layer, uint64_t j) { return 2 * (((j >> 1) << (layer + 1)) + (1ULL << layer) - 1) + (j & 1); }
int gl_tree_prove(gl_ctx* c, gl_handle h, uint64_t leaf_index, uint64_t* out_siblings) {
    TREE(t, h);
    if (leaf_index >= t->n_leaves) return c->fail(GL_ERR_INVALID, "leaf index out of range");
    const unsigned log_sub = log2u(t->n_leaves) - t->cap_height;
    const uint64_t L = 1ULL << log_sub, sub = leaf_index >> log_sub;
    uint64_t j = leaf_index & (L - 1);
    const uint64_t* base = t->digests.data() + 4 * (sub * 2 * (L - 1));
    for (unsigned layer = 0; layer < log_sub; layer++, j >>= 1) std::memcpy(out_siblings + 4 * layer, base + 4 * digest_index(layer, j ^ 1), 32);
    return GL_OK;
}
int gl_tree_open_batch(gl_ctx* c, gl_handle h, const uint64_t* idx, uint32_t n, uint64_t* out_rows, uint64_t* out_siblings) {
    TREE(t, h);
    const unsigned log_sub = log2u(t->n_leaves) - t->cap_height;
    for (uint32_t q = 0; q < n; q++)
        if (idx[q] >= t->n_leaves) return c->fail(GL_ERR_INVALID, "leaf index out of range");
    for (uint32_t q = 0; q < n; q++) {
        if (out_rows) gl_tree_get(c, h, idx[q], out_rows + uint64_t(q) * t->leaf_len);
        if (out_siblings) gl_tree_prove(c, h, idx[q], out_siblings + uint64_t(q) * log_sub * 4);
    }
    return GL_OK;
}
int gl_tree_get_lde_values(gl_ctx* c, gl_handle h, uint64_t index, uint64_t step, uint64_t* out_row) {
    TREE(t, h);
    if (index * step >= t->n_leaves) return c->fail(GL_ERR_INVALID, "index * step out of range");
    return gl_tree_get(c, h, bitrev(index * step, log2u(t->n_leaves)), out_row);
}
int gl_tree_read(gl_ctx* c, gl_handle h, int part, uint64_t* out) {
    TREE(t, h);
    const std::vector<uint64_t>& v = part == GL_PART_COEFFS ? t->coeffs : part == GL_PART_LEAVES ? t->leaves : part == GL_PART_DIGESTS ? t->digests : t->cap;
    if (part == GL_PART_COEFFS && v.empty()) return c->fail(GL_ERR_INVALID, "tree has no coefficients");
    std::memcpy(out, v.data(), 8 * v.size());
    return GL_OK;
}
int gl_tree_free(gl_ctx* c, gl_handle h) {
    if (!c) return GL_ERR_INVALID;
    return c->trees.erase(h) ? GL_OK : GL_ERR_HANDLE;
}

#define FRI(f, h)                                                         \
    auto itf_ = c->fris.find(h);                                          \
    if (itf_ == c->fris.end()) return c->fail(GL_ERR_HANDLE, "unknown handle"); \
    Fri* f = itf_->second.get()

int gl_fri_begin(gl_ctx* c, const uint64_t* coeffs, const uint64_t* values, uint64_t len, uint32_t rate_bits, uint32_t cap_height, gl_handle* out) {
    auto f = std::make_unique<Fri>();
    f->coeffs.assign(coeffs, coeffs + 2 * len); f->values.assign(values, values + 2 * len);
    for (auto& x : f->coeffs) x %= P;
    for (auto& x : f->values) x %= P;
    f->len = len; f->rate_bits = rate_bits; f->cap_height = cap_height;
    *out = c->next; c->fris[c->next++] = std::move(f);
    return GL_OK;
}
int gl_fri_commit_layer(gl_ctx* c, gl_handle h, uint32_t arity_bits, uint64_t* out_leaves, uint64_t* out_digests, uint64_t* out_cap, gl_handle* out_tree) {
    FRI(f, h);
    const unsigned bits = log2u(f->len);
    if (arity_bits > bits) return c->fail(GL_ERR_INVALID, "arity exceeds the layer");
    f->arity_bits = arity_bits;
    std::vector<uint64_t> leaves(2 * f->len);
    for (uint64_t i = 0; i < f->len; i++) { uint64_t s = bitrev(i, bits); leaves[2 * i] = f->values[2 * s]; leaves[2 * i + 1] = f->values[2 * s + 1]; }
    if (out_leaves) std::memcpy(out_leaves, leaves.data(), 16 * f->len);
    return new_tree(c, leaves.data(), f->len >> arity_bits, 2u << arity_bits, f->cap_height, out_digests, out_cap, out_tree);
}
// coeffs <- reduce_with_powers over chunks of the arity committed last; shift <- shift^arity; values <- coset_fft(coeffs, shift)
int gl_fri_fold(gl_ctx* c, gl_handle h, const uint64_t beta_in[2]) {
    FRI(f, h);
    const uint64_t arity = 1ULL << f->arity_bits, m_len = f->len >> f->arity_bits, beta[2] = {beta_in[0] % P, beta_in[1] % P};
    for (uint64_t m = 0; m < m_len; m++) {
        uint64_t acc[2] = {0, 0};
        for (uint64_t t = arity; t-- > 0;) {
            ext_mul(acc, beta, acc);
            acc[0] = add(acc[0], f->coeffs[2 * (arity * m + t)]);
            acc[1] = add(acc[1], f->coeffs[2 * (arity * m + t) + 1]);
        }
        f->coeffs[2 * m] = acc[0]; f->coeffs[2 * m + 1] = acc[1];
    }
    f->len = m_len;
    f->coeffs.resize(2 * m_len); f->values.resize(2 * m_len);
    for (uint32_t k = 0; k < f->arity_bits; k++) f->shift = glo_mul(f->shift, f->shift);
    std::vector<uint64_t> comp(m_len);
    for (int k = 0; k < 2; k++) {
        for (uint64_t i = 0; i < m_len; i++) comp[i] = f->coeffs[2 * i + k];
        glo_coset_fft(comp.data(), log2u(m_len), f->shift);
        for (uint64_t i = 0; i < m_len; i++) f->values[2 * i + k] = comp[i];
    }
    return GL_OK;
}
int gl_fri_final_poly(gl_ctx* c, gl_handle h, uint64_t* out, uint64_t* out_len) {
    FRI(f, h);
    *out_len = f->len >> f->rate_bits;
    if (out) std::memcpy(out, f->coeffs.data(), 16 * *out_len);
    return GL_OK;
}
int gl_fri_end(gl_ctx* c, gl_handle h) { return c->fris.erase(h) ? GL_OK : GL_ERR_HANDLE; }
int gl_fri_read(gl_ctx* c, gl_handle h, uint64_t* out_coeffs, uint64_t* out_values_bitrev, uint64_t* out_len) {
    FRI(f, h);
    if (out_len) *out_len = f->len;
    if (out_coeffs) std::memcpy(out_coeffs, f->coeffs.data(), 16 * f->len);
    if (out_values_bitrev) {
        const unsigned bits = log2u(f->len);
        for (uint64_t i = 0; i < f->len; i++) { uint64_t s = bitrev(i, bits); out_values_bitrev[2 * i] = f->values[2 * s]; out_values_bitrev[2 * i + 1] = f->values[2 * s + 1]; }
    }
    return GL_OK;
}
int gl_fri_pow(gl_ctx*, const uint64_t sponge_state[12], const uint64_t* input_buffer, uint32_t n_inputs, uint32_t min_leading_zeros, uint64_t* out_witness) {
    glo_challenger ch;
    glo_challenger_init(&ch);
    std::memcpy(ch.state, sponge_state, 96);
    for (uint32_t i = 0; i < n_inputs; i++) ch.in[i] = input_buffer[i] % P;
    ch.n_in = int(n_inputs);
    *out_witness = glo_fri_proof_of_work(&ch, min_leading_zeros);
    return GL_OK;
}

#define OPEN(o, h)                                                        \
    auto ito_ = c->openings.find(h);                                      \
    if (ito_ == c->openings.end()) return c->fail(GL_ERR_HANDLE, "unknown handle"); \
    Openings* o = ito_->second.get()

int gl_openings_begin(gl_ctx* c, uint32_t log_n, gl_handle* out) {
    auto o = std::make_unique<Openings>();
    o->log_n = log_n; o->final_poly.assign(2ULL << log_n, 0);
    *out = c->next; c->openings[c->next++] = std::move(o);
    return GL_OK;
}
int gl_openings_add_batch(gl_ctx* c, gl_handle h, const gl_handle* batches, const uint32_t* columns, uint32_t n_polys, const uint64_t alpha[2],
                          const uint64_t point[2], uint64_t* out_quotient) {
    OPEN(o, h);
    const uint64_t n = 1ULL << o->log_n;
    std::vector<const uint64_t*> polys;
    for (uint32_t j = 0; j < n_polys; j++) {
        auto it = c->trees.find(batches[j]);
        if (it == c->trees.end()) return c->fail(GL_ERR_HANDLE, "unknown handle");
        Tree* t = it->second.get();
        if (t->degree_log != o->log_n || t->coeffs.empty()) return c->fail(GL_ERR_INVALID, "Polynomial degrees inconsistent (polynomial %u)", j);
        if (columns[j] >= t->leaf_len) return c->fail(GL_ERR_INVALID, "polynomial %u: column %u out of range", j, columns[j]);
        polys.push_back(t->coeffs.data() + uint64_t(columns[j]) * n);
    }
    const uint64_t a[2] = {alpha[0] % P, alpha[1] % P}, z[2] = {point[0] % P, point[1] % P};
    return glo_openings_add_batch(polys.data(), n_polys, n, a, z, o->final_poly.data(), out_quotient) ? GL_ERR_OOM : GL_OK;
}
int gl_openings_final_poly(gl_ctx* c, gl_handle h, uint64_t* out) {
    OPEN(o, h);
    std::memcpy(out, o->final_poly.data(), 8 * o->final_poly.size());
    return GL_OK;
}
int gl_openings_lde(gl_ctx* c, gl_handle h, uint32_t rate_bits, uint32_t cap_height, gl_handle* out_fri) {
    OPEN(o, h);
    const uint64_t n = 1ULL << o->log_n, R = n << rate_bits;
    std::vector<uint64_t> c0(R, 0), c1(R, 0), lde(2 * R, 0), vals(2 * R);
    for (uint64_t i = 0; i < n; i++) { c0[i] = lde[2 * i] = o->final_poly[2 * i]; c1[i] = lde[2 * i + 1] = o->final_poly[2 * i + 1]; }
    glo_coset_fft(c0.data(), o->log_n + rate_bits, 7);
    glo_coset_fft(c1.data(), o->log_n + rate_bits, 7);
    for (uint64_t i = 0; i < R; i++) { vals[2 * i] = c0[i]; vals[2 * i + 1] = c1[i]; }
    return gl_fri_begin(c, lde.data(), vals.data(), R, rate_bits, cap_height, out_fri);
}
int gl_openings_end(gl_ctx* c, gl_handle h) { return c->openings.erase(h) ? GL_OK : GL_ERR_HANDLE; }
}  // extern "C"
