// merge.cuh -- the agglomerative merge / split initialiser for ONE replay chain.
//
// Reference: blockmodel_t::agg_merge / agg_split / single_block_change / compute_dS(block_move_t) /
// compute_dS(mb, split_move) / compute_b_adj_list / apply_block_moves / apply_split_moves
// (src/blockmodel.cc:109-288, 335-459, 505-611, 639-669), driven by src/mcmc_main.cc:350-451.
//
// Division of labour (the heuristic is a sequential, RNG-driven walk over K x K block data):
//   device  merge_badj_kernel      block adjacency lists (one thread per block, K^2 reads)
//           merge_propose_kernel   the nm x |blist| block-move proposals IN THE REFERENCE'S ORDER, consuming the chain's
//                                  two mt19937 streams (ReplayState) through the libstdc++-exact transforms of
//                                  devmath.cuh; one warp: lane 0 owns the streams and the sequential partial sums of
//                                  discrete_distribution, the 32 lanes stage the weight row and do its divisions
//           merge_dS_kernel        compute_dS of every distinct proposal (one thread per candidate, the reference's
//                                  summation order, lgamma from the glibc table, no FMA contraction): bit-exact
//           split_dS_kernel        compute_dS of every candidate split (one CTA per candidate)
//           merge_first_kernel / merge_relabel_kernel   apply_block_moves: first appearance of every merged block,
//                                  relabelling of the N memberships; init_bisbm = the pool's count builder
//   host    (capi.cu) the priority-queue selection over the <= nm K candidates and the recursion of agg_merge, the
//           geospace ladder (mcmc_main.cc) -- control flow over a few thousand numbers
#pragma once
#include "replay.cuh"

namespace bisbm {

struct MergeCtx {
    ChainRef c;
    Tables tb;
    ReplayState* rs;
    double eps;
    uint32_t na, n;          // nodes of type a, all nodes
    uint32_t* badj;          // [K][stride]: opposite-type blocks mb with m(b, mb) > 0, ascending
    uint32_t* badj_cnt;      // [K]
    uint32_t stride;         // max(ka, kb)
    uint32_t* cand;          // [2 * max_cand]: (source, target) of the distinct proposals, in proposal order
    double* cand_dS;         // [max_cand]
    uint32_t* n_cand;        // [1]
    uint32_t* seen;          // bitmap over (source, target): K * K bits
    uint32_t* first;         // [K] first node carrying each (merged) block
    const uint32_t* map;     // [K] old global block id -> representative (apply) / -> new global id (relabel)
};

#ifdef __CUDACC__

// compute_b_adj_list (src/blockmodel.cc:259-273)
__global__ void merge_badj_kernel(MergeCtx x) {
    const uint32_t K = x.c.ka + x.c.kb;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= K) return;
    const bool ba = b < x.c.ka;
    const uint32_t lo = ba ? x.c.ka : 0u, hi = ba ? K : x.c.ka;   // only blocks of the other type can share edges with b
    uint32_t cnt = 0;
    for (uint32_t mb = lo; mb < hi; ++mb)
        if (m_at(x.c, b, mb) > 0) x.badj[(size_t)b * x.stride + cnt++] = mb;
    x.badj_cnt[b] = cnt;
}

// single_block_change (src/blockmodel.cc:639-669) for src = first_block .. first_block + n_blocks - 1, nm proposals each,
// first occurrence of every (source, target) kept (the reference's set of "source>target" strings).
// One warp; dynamic shared memory: 2 * MT_STATE_WORDS words (the two engines) + K doubles (normalised weights).
__global__ void __launch_bounds__(32) merge_propose_kernel(MergeCtx x, uint32_t first_block, uint32_t n_blocks, uint32_t nm) {
    extern __shared__ __align__(16) unsigned char mg_smem[];
    uint32_t* const engine = reinterpret_cast<uint32_t*>(mg_smem);
    uint32_t* const gen = engine + MT_STATE_WORDS;
    double* const prob = reinterpret_cast<double*>(gen + MT_STATE_WORDS);
    const uint32_t lane = threadIdx.x;
    const ChainRef& c = x.c;
    const uint32_t ka = c.ka, kb = c.kb, K = ka + kb;
    for (uint32_t i = lane; i < MT_STATE_WORDS; i += 32) { engine[i] = x.rs->engine[i]; gen[i] = x.rs->gen[i]; }
    __syncwarp();
    uint32_t ii = 0;
    const double Kd = (double)K;
    for (uint32_t bi = 0; bi < n_blocks; ++bi) {
        const uint32_t src = first_block + bi;
        for (uint32_t rep = 0; rep < nm; ++rep) {
            // lane 0 walks the reference's branches; t != ~0 asks the warp for a categorical draw over row m_[t]
            uint32_t target = src, t = 0xffffffffu;
            bool fixed = (ka == 1 && src < ka) || (kb == 1 && src >= ka);
            if (lane == 0 && !fixed) {
                const uint32_t cnt = x.badj_cnt[src];
                if (cnt == 0) target = (uint32_t)(uint64_t)dmul(mt_canon(engine), Kd);
                else {
                    const uint64_t which = (uint64_t)dmul(mt_canon(engine), (double)cnt);
                    const uint32_t tt = x.badj[(size_t)src * x.stride + which];
                    const double eK = dmul(x.eps, Kd);
                    const double R = ddiv(eK, dadd((double)e_ref(c, slot_of(c, tt)), eK));
                    if (mt_canon(engine) < R) target = (uint32_t)(uint64_t)dmul(mt_canon(engine), Kd);
                    else t = tt;
                }
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t != 0xffffffffu) {
                // std::discrete_distribution<size_t>(m_[t].begin(), m_[t].end())(gen) (libstdc++ bits/random.tcc:2657-2714): the
                // weights are integers, so their sequential double sum is exact and equals the integer sum; the divisions
                // are independent (all lanes); the partial sums are sequential (lane 0)
                // (row t is zero outside the other type's blocks [lo, hi): see rp_categorical for why walking that range alone
                //  returns what lower_bound over the full K-long vector returns)
                const uint32_t lo = t < ka ? ka : 0u, hi = t < ka ? K : ka;
                const double sum = (double)e_ref(c, slot_of(c, t));
                for (uint32_t i = lo + lane; i < hi; i += 32) prob[i] = ddiv((double)m_at(c, t, i), sum);
                __syncwarp();
                if (lane == 0) {
                    if (K < 2) target = 0;
                    else {
                        const double u = mt_canon(gen);
                        double acc = 0.0;
                        target = K - 1;
                        if (!(sum > 0.0) || (lo > 0 && !(0.0 < u))) target = 0;
                        else
                            for (uint32_t i = lo; i < hi; ++i) {
                                acc = (i == 0) ? prob[i] : dadd(acc, prob[i]);
                                const double cp = (i == K - 1) ? 1.0 : acc;
                                if (!(cp < u)) { target = i; break; }   // lower_bound: first cp >= u
                            }
                    }
                }
                __syncwarp();
            }
            if (lane == 0) {
                uint32_t s = src, g = target;
                if (!fixed && !(src > target)) { s = target; g = src; }      // source = the larger id (src == target stays)
                if (fixed) { s = src; g = src; }
                const uint64_t bit = (uint64_t)s * K + g;
                const uint32_t w = x.seen[bit >> 5], m = 1u << (bit & 31u);
                if (!(w & m)) {
                    x.seen[bit >> 5] = w | m;
                    x.cand[2 * ii] = s; x.cand[2 * ii + 1] = g;
                    ++ii;
                }
            }
        }
    }
    __syncwarp();
    for (uint32_t i = lane; i < MT_STATE_WORDS; i += 32) { x.rs->engine[i] = engine[i]; x.rs->gen[i] = gen[i]; }
    if (lane == 0) *x.n_cand = ii;
}

// compute_dS(const block_move_t&) (src/blockmodel.cc:335-370): description-length change of merging block r into s
__global__ void merge_dS_kernel(MergeCtx x, uint32_t n_cand) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const ChainRef& c = x.c;
    const uint32_t ka = c.ka, K = ka + c.kb;
    const uint32_t r = x.cand[2 * i], s = x.cand[2 * i + 1];
    if (r == s || (r < ka && s >= ka) || (r >= ka && s < ka)) { x.cand_dS[i] = BISBM_INF; return; }
    double S0 = 0.0, S1 = 0.0;
    const bool ra = r < ka;
    const uint32_t lo = ra ? ka : 0u, hi = ra ? K : ka;     // criterion(index, KA): the blocks of the other type
    for (uint32_t idx = lo; idx < hi; ++idx) {
        if (e_ref(c, slot_of(c, idx)) == 0) continue;        // _m_r != 0
        const int m_r = m_at(c, r, idx), m_s = m_at(c, s, idx);
        S0 = dsub(S0, lgamma_int(x.tb, (int64_t)m_r + 1));
        S0 = dsub(S0, lgamma_int(x.tb, (int64_t)m_s + 1));
        S1 = dsub(S1, lgamma_int(x.tb, (int64_t)m_s + (int64_t)m_r + 1));
    }
    const int e_r = e_ref(c, slot_of(c, r)), e_s = e_ref(c, slot_of(c, s));
    S0 = dsub(S0, -lgamma_int(x.tb, (int64_t)e_r + 1));
    S0 = dsub(S0, -lgamma_int(x.tb, (int64_t)e_s + 1));
    S1 = dsub(S1, -lgamma_int(x.tb, (int64_t)e_r + (int64_t)e_s + 1));
    x.cand_dS[i] = dsub(S1, S0);
}

// apply_block_moves (src/blockmodel.cc:505-553), first half: map every membership to its set's representative and find the
// first node carrying each representative (the reference renumbers blocks in order of first appearance)
__global__ void merge_first_kernel(MergeCtx x) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= x.n) return;
    const int32_t l = x.c.labels[(size_t)v * x.c.C];
    const uint32_t g = v < x.na ? (uint32_t)l : x.c.ka + (uint32_t)l;
    atomicMin(&x.first[x.map[g]], v);
}
// second half: memberships := new id of the representative (map = old global id -> new global id), stored type-local
// with the new number of type-a blocks
__global__ void merge_relabel_kernel(MergeCtx x, uint32_t new_ka) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= x.n) return;
    const int32_t l = x.c.labels[(size_t)v * x.c.C];
    const uint32_t g = v < x.na ? (uint32_t)l : x.c.ka + (uint32_t)l;
    const uint32_t g2 = x.map[g];
    x.c.labels[(size_t)v * x.c.C] = (int32_t)(v < x.na ? g2 : g2 - new_ka);
}

// compute_dS(size_t mb, vector<bool>& split_move) (src/blockmodel.cc:372-431): description-length change of moving the
// flagged nodes of block r into a new block.  One CTA per candidate; `members` lists the nodes of the candidate's block in
// index order, `flags` (one byte per member) says who moves.  The reference indexes split_move with the running index over
// ALL nodes although the vector only has n_r entries (an out-of-bounds read: undefined behaviour); this kernel implements
// the evident intent -- the index counts the members of the block, as agg_split itself does when it applies the winner
// (src/blockmodel.cc:487-499).  Parity for the split path is therefore "unpinned" (DESIGN.md 4).
struct SplitCand { uint32_t block, first_member, n_members, first_flag; };
__global__ void split_dS_kernel(MergeCtx x, GraphView G, const SplitCand* cands, const uint32_t* members, const uint8_t* flags,
                                double* out) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    int* const k = reinterpret_cast<int*>(sp_smem);           // [kopp] edges from the moving nodes to each block of the other type
    __shared__ int s_deg;
    const ChainRef& c = x.c;
    const SplitCand cd = cands[blockIdx.x];
    const uint32_t ka = c.ka, K = ka + c.kb, r = cd.block;
    const bool ra = r < ka;
    const uint32_t kopp = ra ? c.kb : ka;
    for (uint32_t i = threadIdx.x; i < kopp; i += blockDim.x) k[i] = 0;
    if (threadIdx.x == 0) s_deg = 0;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < cd.n_members; j += blockDim.x) {
        if (!flags[cd.first_flag + j]) continue;
        const uint32_t v = members[cd.first_member + j];
        const uint32_t e0 = G.row_ptr[v], e1 = G.row_ptr[v + 1];
        for (uint32_t e = e0; e < e1; ++e) atomicAdd(&k[c.labels[(size_t)G.col[e] * c.C]], 1);
        atomicAdd(&s_deg, (int)(e1 - e0));
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (cd.n_members == 0) { out[blockIdx.x] = BISBM_INF; return; }
    double S0 = 0.0, S1 = 0.0;
    const uint32_t lo = ra ? ka : 0u, hi = ra ? K : ka;
    for (uint32_t idx = lo; idx < hi; ++idx) {
        const int m_r = m_at(c, r, idx), kk = k[idx - lo];
        S0 = dsub(S0, lgamma_int(x.tb, (int64_t)m_r + 1));
        S1 = dsub(S1, lgamma_int(x.tb, (int64_t)m_r - kk + 1));
        S1 = dsub(S1, lgamma_int(x.tb, (int64_t)kk + 1));
    }
    const int e_r = e_ref(c, slot_of(c, r)), deg = s_deg;
    S0 = dsub(S0, -lgamma_int(x.tb, (int64_t)e_r + 1));
    S1 = dsub(S1, -lgamma_int(x.tb, (int64_t)e_r - deg + 1));
    S1 = dsub(S1, -lgamma_int(x.tb, (int64_t)deg + 1));
    out[blockIdx.x] = dsub(S1, S0);
}

#endif  // __CUDACC__

}  // namespace bisbm
