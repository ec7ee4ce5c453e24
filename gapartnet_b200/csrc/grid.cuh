// Occupancy directory of one sparse level ("bitmap-rank perfect hash").
//
// A level with B scenes on an X*Y*Z grid keeps one bit per cell (z fastest,
// scenes padded to a multiple of 32 cells) plus an exclusive popcount prefix
// per 32-bit word.  row(cell) = prefix[word] + popc(bits below) is collision
// free, O(1), needs no key compare, and enumerates the active cells in
// lexicographic (b,x,y,z) order - the canonical row order of this library.
// It replaces the coordinate hash table of spconv's indice-pair builder
// (reference call sites: gapartnet/network/backbone.py:25-28,74-77,87-90).
#pragma once
#include "common.cuh"

struct GridDir {
    const uint32_t* words;    // [n_words]
    const int* prefix;        // [n_words + 1] exclusive popcount prefix
    const int* row_of_rank;   // optional [M]: rank -> caller row (NULL = identity)
    int B, X, Y, Z;
    uint32_t scene_stride;    // roundup(X*Y*Z, 32)
};

static inline uint32_t gp_scene_stride(int X, int Y, int Z) {
    unsigned long long c = (unsigned long long)X * Y * Z;
    return (uint32_t)((c + 31ull) & ~31ull);
}
static inline long long gp_grid_words(int B, int X, int Y, int Z) {
    return (long long)B * gp_scene_stride(X, Y, Z) / 32;
}

__device__ __forceinline__ uint32_t grid_cell(const GridDir& g, int b, int x, int y, int z) {
    return (uint32_t)b * g.scene_stride + (uint32_t)((x * g.Y + y) * g.Z + z);
}

// rank of an occupied cell (caller guarantees the bit is set)
__device__ __forceinline__ int grid_rank(const GridDir& g, uint32_t cell) {
    uint32_t w = cell >> 5, bit = cell & 31u;
    uint32_t word = __ldg(g.words + w);
    return __ldg(g.prefix + w) + __popc(word & ((1u << bit) - 1u));
}

// row id of (b,x,y,z) or -1 when out of the grid / unoccupied
__device__ __forceinline__ int grid_lookup(const GridDir& g, int b, int x, int y, int z) {
    if ((unsigned)x >= (unsigned)g.X || (unsigned)y >= (unsigned)g.Y ||
        (unsigned)z >= (unsigned)g.Z)
        return -1;
    uint32_t cell = grid_cell(g, b, x, y, z);
    uint32_t w = cell >> 5, bit = cell & 31u;
    uint32_t word = __ldg(g.words + w);
    if (!((word >> bit) & 1u)) return -1;
    int r = __ldg(g.prefix + w) + __popc(word & ((1u << bit) - 1u));
    return g.row_of_rank ? __ldg(g.row_of_rank + r) : r;
}

// host-side launchers implemented in grid.cu -------------------------------
// exclusive popcount scan of words[0..n_words) -> prefix[0..n_words]; the total
// goes to prefix[n_words] and (if non-NULL) *d_total.  scan_tmp: >= n_words/2048 + 2 ints.
int gp_grid_scan(const uint32_t* words, long long n_words, int* prefix, int* scan_tmp,
                 int* d_total, cudaStream_t stream);
