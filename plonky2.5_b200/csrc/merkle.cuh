// K4/K5: Poseidon Merkle tree in plonky2's interleaved digest layout.
//
// Replaces plonky2 hash/src/merkle_tree.rs · MerkleTree::new / fill_digests_buf / fill_subtree and
// plonk/config.rs · Hasher::hash_or_noop, hash/hashing.rs · compress (semantics SURVEY.md A.5, A.6).
// Driven in the reference from /root/reference/src/p3/mod.rs:250 and :260 (every PolynomialBatch commit and every
// FRI commit-phase layer builds one tree).
//
// Layout: leaves are a row-major device matrix [n_leaves][pitch] (pitch multiple of 8 words = 64 B);
// digests = 2*(n_leaves - 2^cap_height) HashOuts; subtree t occupies [t*2(L-1), (t+1)*2(L-1)); inside a subtree the
// node j of layer i (0 = leaf digests) sits at 2*(((j>>1) << (i+1)) + (1<<i) - 1) + (j&1); roots go to cap[t].
#pragma once
#include "poseidon.cuh"

namespace merkle {

__host__ __device__ __forceinline__ uint64_t digest_index(uint32_t layer, uint64_t j) {
    return 2 * (((j >> 1) << (layer + 1)) + (1ULL << layer) - 1) + (j & 1);
}

__device__ __forceinline__ void store_digest(uint64_t* dst, const uint64_t (&s)[poseidon::WIDTH]) {
    ulonglong2* d = reinterpret_cast<ulonglong2*>(dst);
    d[0] = make_ulonglong2(gl::canon(s[0]), gl::canon(s[1]));
    d[1] = make_ulonglong2(gl::canon(s[2]), gl::canon(s[3]));
}

// One thread per leaf: hash_or_noop(row) with the overwrite-mode rate-8 sponge.
#ifndef LEAF_MIN_BLOCKS
#define LEAF_MIN_BLOCKS 6   // 80 registers: measured 43.2 ms vs 44.0 ms at 94-106 registers (2^22 x 135 leaves)
#endif
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, LEAF_MIN_BLOCKS)
leaf_hash_kernel(const uint64_t* __restrict__ leaves, uint32_t pitch, uint32_t leaf_len, uint64_t n_leaves,
                 uint32_t log_sub, uint64_t* __restrict__ digests, uint64_t* __restrict__ cap) {
    uint64_t leaf = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (leaf >= n_leaves) return;
    const uint64_t* row = leaves + leaf * pitch;
    uint64_t s[poseidon::WIDTH];
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) s[i] = 0;
    if (leaf_len <= 4) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if ((uint32_t)i < leaf_len) s[i] = row[i];   // hash_or_noop: no permutation
    } else {
        const uint32_t n_chunks = (leaf_len + 7) >> 3;
#pragma unroll 1
        for (uint32_t k = 0; k < n_chunks; k++) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(row + 8 * k);
            ulonglong2 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
            uint64_t in[8] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y};
            const uint32_t rem = leaf_len - 8 * k;   // >= 1
#pragma unroll
            for (int i = 0; i < 8; i++)
                if ((uint32_t)i < rem) s[i] = in[i];   // overwrite mode: a short last chunk keeps the other lanes
            poseidon::permute(s);
        }
    }
    uint64_t* dst;
    if (log_sub == 0) dst = cap + 4 * leaf;
    else {
        uint64_t L = 1ULL << log_sub, t = leaf >> log_sub, j = leaf & (L - 1);
        dst = digests + 4 * (t * 2 * (L - 1) + digest_index(0, j));
    }
    store_digest(dst, s);
}

// Streaming form of the leaf sponge: absorbs columns [col0, col1) of every leaf (col0 a multiple of the rate 8; col1 a multiple of 8 or the
// leaf's end) and carries the 12-word sponge state in `state` ([12][n_leaves], one coalesced word per lane and leaf) between launches.
// The overwrite-mode sponge consumes a row strictly left to right, so the column chunks of a host->device commit can be hashed as soon as
// their LDE is done, while later chunks are still crossing PCIe: the copy hides behind the hashing instead of sitting in front of it.
// first: start from the zero state; last: write the digest instead of the state.  Same permutations, same order => same digests.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, LEAF_MIN_BLOCKS)
leaf_absorb_kernel(const uint64_t* __restrict__ leaves, uint32_t pitch, uint32_t leaf_len, uint32_t col0, uint32_t col1, uint64_t n_leaves,
                   uint32_t log_sub, uint64_t* __restrict__ state, uint64_t* __restrict__ digests, uint64_t* __restrict__ cap, int first, int last) {
    const uint64_t leaf = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (leaf >= n_leaves) return;
    const uint64_t* row = leaves + leaf * pitch;
    uint64_t s[poseidon::WIDTH];
    // overwrite mode: a full 8-column group replaces lanes 0..7, so only the capacity lanes 8..11 have to travel between launches
    // unless the group that follows the boundary is the leaf's short tail (which keeps the lanes it does not cover)
    const bool in_full = leaf_len - col0 >= 8;
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) s[i] = (first || (i < 8 && in_full)) ? 0 : state[(uint64_t)i * n_leaves + leaf];
#pragma unroll 1
    for (uint32_t k = col0; k < col1; k += 8) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(row + k);
        ulonglong2 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
        uint64_t in[8] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y, v3.x, v3.y};
        const uint32_t rem = leaf_len - k;   // >= 1
#pragma unroll
        for (int i = 0; i < 8; i++)
            if ((uint32_t)i < rem) s[i] = in[i];
        poseidon::permute(s);
    }
    if (!last) {
        const bool out_full = leaf_len - col1 >= 8;
#pragma unroll
        for (int i = 0; i < poseidon::WIDTH; i++)
            if (i >= 8 || !out_full) state[(uint64_t)i * n_leaves + leaf] = s[i];
        return;
    }
    uint64_t* dst;
    if (log_sub == 0) dst = cap + 4 * leaf;
    else {
        uint64_t L = 1ULL << log_sub, t = leaf >> log_sub, j = leaf & (L - 1);
        dst = digests + 4 * (t * 2 * (L - 1) + digest_index(0, j));
    }
    store_digest(dst, s);
}

// One thread per node of layer `layer` (1..log_sub): two_to_one(children).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, LEAF_MIN_BLOCKS)
tree_level_kernel(uint64_t* __restrict__ digests, uint64_t* __restrict__ cap, uint32_t layer, uint32_t log_sub,
                  uint64_t n_nodes) {
    uint64_t node = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (node >= n_nodes) return;
    const uint32_t lw = log_sub - layer;                 // log2(nodes of this layer per subtree)
    const uint64_t t = node >> lw, j = node & ((1ULL << lw) - 1);
    const uint64_t L = 1ULL << log_sub;
    uint64_t* base = digests + 4 * (t * 2 * (L - 1));
    const ulonglong2* ch = reinterpret_cast<const ulonglong2*>(base + 4 * digest_index(layer - 1, 2 * j));
    ulonglong2 a = ch[0], b = ch[1], c = ch[2], d = ch[3];   // left digest then right digest (adjacent)
    uint64_t s[poseidon::WIDTH] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, 0, 0, 0, 0};
    poseidon::permute(s);
    uint64_t* dst = (layer == log_sub) ? cap + 4 * t : base + 4 * digest_index(layer, j);
    store_digest(dst, s);
}

// FRI query rounds (plonky2 fri/prover.rs · fri_prover_query_round): MerkleTree::get + MerkleTree::prove for a batch of leaf
// indices in one launch.  Thread (q, w): word w of the answer for query q — first the leaf row, then depth*4 sibling words.
__global__ void open_batch_kernel(const uint64_t* __restrict__ leaves, uint32_t pitch, uint32_t leaf_len, const uint64_t* __restrict__ digests,
                                  uint32_t log_sub, const uint64_t* __restrict__ indices, uint32_t n_queries,
                                  uint64_t* __restrict__ out_rows, uint64_t* __restrict__ out_siblings) {
    const uint32_t per_q = leaf_len + 4 * log_sub;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n_queries * per_q) return;
    const uint32_t q = (uint32_t)(i / per_q), w = (uint32_t)(i % per_q);
    const uint64_t leaf = indices[q];
    if (w < leaf_len) {
        if (out_rows) out_rows[(size_t)q * leaf_len + w] = leaves[leaf * pitch + w];
        return;
    }
    if (!out_siblings) return;
    const uint32_t layer = (w - leaf_len) >> 2, k = (w - leaf_len) & 3;
    const uint64_t L = 1ULL << log_sub, sub = leaf >> log_sub, j = (leaf & (L - 1)) >> layer;
    out_siblings[((size_t)q * log_sub + layer) * 4 + k] = digests[4 * (sub * 2 * (L - 1) + digest_index(layer, j ^ 1)) + k];
}

// Batch of independent permutations (host challenger / PoW plumbing): states[n][12] in place, canonical out.
__global__ void permute_kernel(uint64_t* __restrict__ states, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s[poseidon::WIDTH];
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) s[k] = states[12 * i + k];
    poseidon::permute(s);
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) states[12 * i + k] = gl::canon(s[k]);
}

// The Challenger's duplexing loop for a run of FULL input groups (plonky2 iop/challenger.rs · observe_elements -> duplexing, overwrite
// mode): state[0..8) <- group k, permute, for k = 0..n_groups.  The chain is strictly sequential (each permutation needs the previous
// state), so one thread runs it: a single launch and one round trip for a whole Merkle cap (8 groups) instead of one per permutation.
__global__ void sponge_absorb_kernel(uint64_t* __restrict__ state12, const uint64_t* __restrict__ groups, uint32_t n_groups) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t s[poseidon::WIDTH];
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) s[k] = state12[k];
#pragma unroll 1
    for (uint32_t g = 0; g < n_groups; g++) {
#pragma unroll
        for (int k = 0; k < poseidon::RATE; k++) s[k] = groups[8 * g + k];
        poseidon::permute(s);
    }
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) state12[k] = gl::canon(s[k]);
}

// K8: FRI proof-of-work grinding (plonky2 fri/prover.rs · fri_proof_of_work).  Candidate w goes to lane
// `pos` of the pre-absorbed sponge state, one permutation, accept when canonical state[7] has >= min_lz leading zero
// bits.  One thread per candidate of the window [base, base + n); the smallest accepted candidate wins (atomicMin), which
// is the serial `find` the reference runs (plonky2 `parallel` off) and keeps the proof deterministic.
__global__ void pow_grind_kernel(const uint64_t* __restrict__ state12, uint32_t pos, uint32_t min_lz, uint64_t base, uint64_t n,
                                 unsigned long long* __restrict__ best) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t w = base + i;
    if (w >= gl::P) return;                      // candidates are canonical field elements
    uint64_t s[poseidon::WIDTH];
#pragma unroll
    for (int k = 0; k < poseidon::WIDTH; k++) s[k] = (uint32_t)k == pos ? w : state12[k];
    poseidon::permute(s);
    const uint64_t r = gl::canon(s[poseidon::RATE - 1]);
    const uint32_t lz = r ? (uint32_t)__clzll((long long)r) : 64u;
    if (lz >= min_lz) atomicMin(best, (unsigned long long)w);
}

}  // namespace merkle
