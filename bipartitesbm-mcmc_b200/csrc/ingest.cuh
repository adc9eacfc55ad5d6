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

}  // namespace bisbm
