// The recorded witness program (b2r_prog): device-resident value graph + cell map, and the host copies
// of the fixed columns, range tags and copy constraints that keygen consumes (prover.cu).
#pragma once
#include <array>
#include <vector>

#include "circuit.hpp"
#include "field.cuh"

using b2r::fe_t;
using namespace b2r::circuit;

struct LevelRange {
    uint32_t start, end;    // node ids
    uint32_t bstart, bend;  // range in big_nodes
};

struct b2r_prog {
    uint32_t bits_len = 0, num_limbs = 0, k = 0;
    uint32_t rows_used = 0, num_values = 0, num_levels = 0;
    int32_t is_valid_vid = -1;
    uint32_t num_inputs = 0;
    uint32_t aux_words = 4;  // 64-bit words per instance in the third input array: 4 hash limbs (+ the exponent for RSAPubE::Var)
    // device
    Node* d_nodes = nullptr;
    LevelRange* d_levels = nullptr;
    uint32_t* d_big_nodes = nullptr;
    BigOp* d_big_ops = nullptr;
    uint32_t* d_big_inputs = nullptr;
    fe_t* d_consts = nullptr;
    int32_t* d_cellmap = nullptr;  // [5][2^k]
    uint32_t max_big_words = 0;
    // host copies kept for keygen / inspection
    std::vector<std::array<uint32_t, NUM_FIXED>> fixed;
    std::vector<std::array<uint8_t, 4>> range_tags;
    std::vector<std::array<uint32_t, 4>> copies;
    std::vector<U256> constants;
    uint8_t tag_bits[16] = {0};  // bits of lookup tag t (RangeChip table), 0 = unused
};

