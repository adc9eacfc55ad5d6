// ingest.cuh -- edge list -> device-resident CSR, on the device.
//
// Replaces the host-side counting sort of round 1 for edge_to_adj (reference src/graph_utilities.cc:36-49): the
// adjacency of node v lists its neighbours in FILE ORDER, both directions pushed, multi-edges kept.  Directed entry
// 2i is (a_i -> b_i), entry 2i+1 is (b_i -> a_i); a STABLE radix sort by source node keeps each row in file order.
// The sort itself is CUB's DeviceRadixSort (a library call on a one-off, non-hot path); everything else is small
// kernels.  The label-independent entropy terms (src/blockmodel.cc:755-757, 772-779) come from two integer
// histograms (degrees, multi-edge multiplicities) so that they are summed on the host in a fixed order.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdint.h>

namespace bisbm {

// codes written to bad[1] with the offending edge index in bad[0]
enum { INGEST_BAD_RANGE = 1, INGEST_BAD_TYPE = 2 };

__global__ void ingest_expand_kernel(const uint32_t* __restrict__ ea, const uint32_t* __restrict__ eb, uint64_t n_edges, uint32_t n,
                                     uint32_t na, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const uint32_t a = ea[i], b = eb[i];
    if (a >= n || b >= n) { atomicMin(&bad[0], (unsigned long long)i * 4 + INGEST_BAD_RANGE); keys[2 * i] = 0; vals[2 * i] = 0; keys[2 * i + 1] = 0; vals[2 * i + 1] = 0; return; }
    if ((a < na) == (b < na)) atomicMin(&bad[0], (unsigned long long)i * 4 + INGEST_BAD_TYPE);
    keys[2 * i] = a; vals[2 * i] = b;
    keys[2 * i + 1] = b; vals[2 * i + 1] = a;
}

__global__ void ingest_degree_kernel(const uint32_t* __restrict__ keys, uint64_t n_entries, uint32_t* __restrict__ deg) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_entries) atomicAdd(&deg[keys[p]], 1u);
}

__global__ void ingest_degidx_kernel(const uint32_t* __restrict__ deg, uint32_t n, const uint32_t* __restrict__ table, uint32_t* __restrict__ degidx) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) degidx[v] = table[deg[v]];
}

// one 64-bit key (larger id << 32 | smaller id) per undirected edge: equal keys = a multi-edge
__global__ void ingest_pair_keys_kernel(const uint32_t* __restrict__ ea, const uint32_t* __restrict__ eb, uint64_t n_edges,
                                        unsigned long long* __restrict__ keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const uint32_t a = ea[i], b = eb[i];
    keys[i] = a > b ? ((unsigned long long)a << 32) | b : ((unsigned long long)b << 32) | a;
}

enum { INGEST_MULT_BINS = 4096 };
// mult[L] += 1 for every run of L equal keys (L >= 2; L >= INGEST_MULT_BINS lands in the last bin)
__global__ void ingest_multiplicity_kernel(const unsigned long long* __restrict__ keys, uint64_t n_edges, unsigned long long* __restrict__ mult) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_edges) return;
    if (p != 0 && keys[p - 1] == keys[p]) return;          // not a run start
    uint64_t L = 1;
    while (p + L < n_edges && keys[p + L] == keys[p]) ++L;
    if (L >= 2) atomicAdd(&mult[L < INGEST_MULT_BINS ? L : INGEST_MULT_BINS - 1], 1ull);
}

// ---- edge-list TEXT -> edge arrays, on the device (load_edge_list, reference src/graph_utilities.cc:20-34) -------------
// One line per edge, two unsigned integers separated by blanks (`linestream >> node_a >> node_b`); lines whose first token is
// not a number are skipped, a missing second number reads as 0.  Two passes over the text with the same kernel: the first
// counts the edge lines of every 4 KB chunk, an exclusive scan turns the counts into output offsets, the second parses each
// line at its start and writes edge i of the FILE ORDER to slot i (a line belongs to the chunk its first byte is in; the
// parse may read past the chunk's end).  16 bytes per thread, one vector load.
enum { PARSE_T = 256, PARSE_B = 16, PARSE_CHUNK = PARSE_T * PARSE_B };

__device__ __forceinline__ bool parse_blank(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
__device__ __forceinline__ bool parse_digit(unsigned char c) { return c >= '0' && c <= '9'; }

template <bool WRITE>
__global__ void __launch_bounds__(PARSE_T) parse_edges_kernel(const unsigned char* __restrict__ text, uint64_t bytes,
                                                               unsigned long long* __restrict__ chunk_lines,   // !WRITE: out, lines per chunk; WRITE: in, first slot of the chunk
                                                               uint32_t* __restrict__ ea, uint32_t* __restrict__ eb, unsigned long long* bad) {
    __shared__ uint32_t warp_tot[PARSE_T / 32];
    const uint64_t base = (uint64_t)blockIdx.x * PARSE_CHUNK + (uint64_t)threadIdx.x * PARSE_B;
    unsigned char c[PARSE_B];
    if (base + PARSE_B <= bytes) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + base);       // (the buffer is 16-byte aligned and padded)
        memcpy(c, &v, PARSE_B);
    } else {
        for (int k = 0; k < PARSE_B; ++k) c[k] = (base + k < bytes) ? text[base + k] : (unsigned char)'\n';
    }
    unsigned char prev = (base == 0) ? (unsigned char)'\n' : (base - 1 < bytes ? text[base - 1] : (unsigned char)'\n');
    uint32_t starts = 0;
    for (int k = 0; k < PARSE_B; ++k) {
        if (prev == '\n' && base + k < bytes) {
            uint64_t j = base + k;                                   // an edge line: blanks, then a digit
            while (j < bytes && parse_blank(text[j])) ++j;
            if (j < bytes && parse_digit(text[j])) starts |= 1u << k;
        }
        prev = c[k];
    }
    // exclusive scan of the per-thread line counts over the block
    const uint32_t cnt = (uint32_t)__popc(starts), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += y; }
    if (lane == 31u) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t w = 0; w < PARSE_T / 32; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
    if (!WRITE) {
        if (threadIdx.x == 0) chunk_lines[blockIdx.x] = total;
        return;
    }
    uint64_t slot = chunk_lines[blockIdx.x] + before + (incl - cnt);
    while (starts) {
        const int k = __ffs((int)starts) - 1;
        starts &= starts - 1;
        uint64_t j = base + k;
        while (j < bytes && parse_blank(text[j])) ++j;
        unsigned long long a = 0, b = 0;
        bool over = false;
        while (j < bytes && parse_digit(text[j])) { a = a * 10ull + (text[j] - '0'); over |= a > 0xffffffffull; ++j; }
        while (j < bytes && parse_blank(text[j])) ++j;
        while (j < bytes && parse_digit(text[j])) { b = b * 10ull + (text[j] - '0'); over |= b > 0xffffffffull; ++j; }
        if (over) atomicMin(bad, (unsigned long long)slot);
        ea[slot] = (uint32_t)a; eb[slot] = (uint32_t)b;
        ++slot;
    }
}

}  // namespace bisbm
