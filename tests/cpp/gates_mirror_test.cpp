// The C++ mirror's gate / quotient / permutation entry points (include/gl_plonky2.hpp) on the GPU: self-checking identities only
// (the word-for-word comparison with the oracle lives in tests/test_gates.py):
//   * rows filled by poseidon2_gate_witness satisfy every Poseidon2Gate constraint;
//   * a wires batch made of such rows (constant columns => every LDE row equals the row) accumulates a zero quotient, and the
//     quotient commit of the zero polynomial has all-zero coefficients;
//   * with sigma = identity permutation the partial products are all 1 (numerator = denominator) and Z = 1.
// Build: g++ -std=c++17 -O2 -I include tests/cpp/gates_mirror_test.cpp -o tests/cpp/build/gates_mirror_test -L plonky2.5_b200 -lgl_commit
#include <cstdio>
#include <cstdlib>

#include "gl_plonky2.hpp"

using namespace plonky2;

static int failures = 0;
#define CHECK(cond, msg) do { if (!(cond)) { std::printf("FAIL: %s\n", msg); failures++; } } while (0)

int main() {
    try {
        Context ctx(0);
        const uint64_t P = 0xFFFFFFFF00000001ULL;
        // (1) witness rows satisfy the gate
        std::vector<F> in;
        uint64_t z = 12345;
        for (int r = 0; r < 8; r++) {
            for (int i = 0; i < 12; i++) { z = z * 6364136223846793005ULL + 1442695040888963407ULL; in.push_back(z % P); }
            in.push_back(uint64_t(r & 1));
        }
        std::vector<F> rows = poseidon2_gate_witness(in, &ctx);
        CHECK(rows.size() == 8 * 135, "witness size");
        std::vector<F> cons = evaluate_gate_constraints(GL_GATE_POSEIDON2, 0, rows, &ctx);
        bool zero = cons.size() == 8 * 123;
        for (F v : cons) zero = zero && v == 0;
        CHECK(zero, "generated rows satisfy all 123 constraints");
        rows[50] ^= 1;
        cons = evaluate_gate_constraints(GL_GATE_POSEIDON2, 0, rows, &ctx);
        bool any = false;
        for (F v : cons) any = any || v != 0;
        CHECK(any, "a tampered wire breaks a constraint");
        rows[50] ^= 1;
        // (2) constant columns: every LDE row is row 0 of the witness => zero quotient
        const size_t n = 16;
        std::vector<PolynomialValues> wires(135);
        for (size_t j = 0; j < 135; j++) wires[j].values.assign(n, rows[j]);
        PolynomialBatch batch = PolynomialBatch::from_values(wires, 3, false, 2, nullptr, nullptr, &ctx);
        QuotientAccumulator q(batch, 2, &ctx);
        q.add_gate(GL_GATE_POSEIDON2, 0, {7, 11});
        std::vector<F> acc = q.values();
        zero = acc.size() == 2 * n * 8;
        for (F v : acc) zero = zero && v == 0;
        CHECK(zero, "valid rows accumulate a zero quotient");
        PolynomialBatch qb = q.commit(2);
        CHECK(qb.merkle_tree.cap.len() == 4 && qb.degree_log == 4, "quotient commit shape");
        // (3) identity permutation: sigma_j(x) = k_j x  =>  every chunk quotient is 1
        const size_t nr = 10, log_n = 4;
        std::vector<F> k_is(nr), xs(n);
        uint64_t g = 1;
        for (size_t j = 0; j < nr; j++) { k_is[j] = g; g = (unsigned __int128)g * 7 % P; }
        uint64_t w = 1753635133440165772ULL;
        for (size_t i = 0; i < 32 - log_n; i++) w = (unsigned __int128)w * w % P;
        uint64_t x = 1;
        for (size_t i = 0; i < n; i++) { xs[i] = x; x = (unsigned __int128)x * w % P; }
        std::vector<PolynomialValues> wv(nr), sv(nr);
        for (size_t j = 0; j < nr; j++)
            for (size_t i = 0; i < n; i++) {
                wv[j].values.push_back((i * 31 + j * 17 + 5) % P);
                sv[j].values.push_back((unsigned __int128)k_is[j] * xs[i] % P);
            }
        auto pp = partial_products_and_zs(wv, sv, k_is, {3, 5}, {9, 13}, 4, &ctx);
        bool ones = pp.size() == 2 * 3;
        for (const auto& poly : pp)
            for (F v : poly.values) ones = ones && v == 1;
        CHECK(ones, "identity permutation: Z and all partial products are 1");
    } catch (const std::exception& e) {
        std::printf("FAIL: exception: %s\n", e.what());
        failures++;
    }
    std::printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
