// state.cuh -- device-resident block-model state (SoA), replacing blockmodel_t's host
// std::vectors (reference src/blockmodel.hh:92-143).
//
// Layout in HBM (C = number of chains padded to a multiple of 32, KA/KB = maxima over chains):
//   graph   row_ptr u32[n+1], col u32[2E] (reference adjacency order, multi-edges kept),
//           degidx u32[n] (index of deg(v) in the sorted list of distinct degrees, width W)
//   labels  i32[n][C]   CHAIN-MINOR, type-local block index (type-a node: 0..ka-1, type-b: 0..kb-1)
//   counts are GROUP-INTERLEAVED: chain c lives in group c/32 at lane c%32, and the 32 chains of a
//   group are the fastest index, so lane = chain accesses of one entry are one 128-byte line and
//   a group's counts copy straight into bank-conflict-free shared memory:
//   m       i32[C/32][KA][KB][32]   inter-type edge counts m_rs, r in type a, s in type b
//           (the reference keeps the full symmetric K x K matrix; only this block is non-zero)
//   e       i32[C/32][KA+KB][32]    block degree totals e_r (reference m_r_): type-a at [0,KA), type-b at [KA,KA+KB)
//   nr      i32[C/32][KA+KB][32]    block sizes n_r
//   eta     i32[C/32][KA+KB][W][32] number of nodes of block r with the w-th distinct degree (reference eta_rk_)
// The reference's N x K neighbour-block matrix k_ is NOT stored: the histogram of a vertex is
// rebuilt from its neighbours' labels on every move (that gather is the HBM stream).
#pragma once
#include <stdint.h>
#include "devmath.cuh"

namespace bisbm {

struct GraphView {
    uint32_t n, na, nb;
    uint64_t n_edges;
    const uint32_t* row_ptr;
    const uint32_t* col;
    const uint32_t* degidx;
    uint32_t W;  // number of distinct degrees
    uint32_t max_degree;
};

struct StateView {
    uint32_t C;   // padded chain count (label stride)
    uint32_t KA, KB;  // maxima (strides)
    uint32_t W;
    int32_t* labels;
    int32_t* m;
    int32_t* e;
    int32_t* nr;
    int32_t* eta;
    const uint32_t* ka;  // [C]
    const uint32_t* kb;  // [C]
    double eps;
};

// accessors for one chain
struct ChainRef {
    int32_t* labels;  // + chain offset; index v*C
    uint32_t C;
    int32_t* m;
    int32_t* e;
    int32_t* nr;
    int32_t* eta;
    uint32_t ka, kb, KA, KB, W;
    uint32_t cs;  // stride between consecutive count entries of this chain (32 on the device)
};

enum { GROUP = 32 };

// offset of chain c's first entry in a group-interleaved array with `per_chain` entries per chain
BISBM_HD size_t cnt_base(uint32_t c, size_t per_chain) { return (size_t)(c / GROUP) * per_chain * GROUP + (c % GROUP); }

// ka / kb are passed in (host copies) so this also works on the host side
BISBM_HD ChainRef chain_ref(const StateView& s, uint32_t c, uint32_t ka, uint32_t kb) {
    ChainRef r;
    r.labels = s.labels + c;
    r.C = s.C;
    r.m = s.m + cnt_base(c, (size_t)s.KA * s.KB);
    r.e = s.e + cnt_base(c, (size_t)s.KA + s.KB);
    r.nr = s.nr + cnt_base(c, (size_t)s.KA + s.KB);
    r.eta = s.eta + cnt_base(c, ((size_t)s.KA + s.KB) * s.W);
    r.ka = ka; r.kb = kb;
    r.KA = s.KA; r.KB = s.KB; r.W = s.W;
    r.cs = GROUP;
    return r;
}

// slot of global block id g in the e / nr / eta arrays
BISBM_HD uint32_t slot_of(const ChainRef& c, uint32_t g) { return g < c.ka ? g : c.KA + (g - c.ka); }
// m entry between global block ids g and h (0 when both are of the same type)
BISBM_HD int32_t m_at(const ChainRef& c, uint32_t g, uint32_t h) {
    bool ga = g < c.ka, ha = h < c.ka;
    if (ga == ha) return 0;
    uint32_t a = ga ? g : h, b = ga ? h - c.ka : g - c.ka;
    return c.m[((size_t)a * c.KB + b) * c.cs];
}
BISBM_HD int32_t* m_ptr(const ChainRef& c, uint32_t g, uint32_t h) {  // g, h of different types
    bool ga = g < c.ka;
    uint32_t a = ga ? g : h, b = ga ? h - c.ka : g - c.ka;
    return c.m + ((size_t)a * c.KB + b) * c.cs;
}
BISBM_HD int32_t& e_ref(const ChainRef& c, uint32_t slot) { return c.e[(size_t)slot * c.cs]; }
BISBM_HD int32_t& nr_ref(const ChainRef& c, uint32_t slot) { return c.nr[(size_t)slot * c.cs]; }
BISBM_HD int32_t& eta_ref(const ChainRef& c, uint32_t slot, uint32_t w) { return c.eta[((size_t)slot * c.W + w) * c.cs]; }

}  // namespace bisbm
