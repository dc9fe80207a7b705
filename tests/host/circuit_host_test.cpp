// Host-side exerciser for the circuit recorder (csrc/circuit.hpp: the mirror of the reference's
// RSAChip / BigIntChip / maingate interface).  Records the pkcs1v15 bench circuit and prints a
// layout digest that tests/test_host_circuit.py compares with the oracle's independently built
// table (oracle/rsa_witness.c: orc_table_layout_digest).  No GPU, no values: layout only.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../../halo2-rsa_b200/csrc/circuit.hpp"
using namespace b2r::circuit;

static uint64_t mix64(uint64_t h, uint64_t v) { h ^= v; h *= 0x100000001B3ull; h ^= h >> 29; return h; }

// usage: circuit_host_test <bits> <k> <e>                 pkcs1v15 circuit, fixed exponent
//        circuit_host_test <bits> <k> var <exp_limb_bits>   pkcs1v15 circuit, RSAPubE::Var
//        circuit_host_test <bits> <k> op <id> <exp_limb_bits>   single BigIntChip operation (BigIntTestOp ids)
//        circuit_host_test aux <limb_width> <nl> <nr>        RefreshAux::new(...).increased_limbs_vec
int main(int argc, char** argv) {
    if (argc > 4 && std::string(argv[1]) == "aux") {
        auto inc = BigIntChip::refresh_increased_limbs(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
        printf("aux=");
        for (size_t i = 0; i < inc.size(); i++) printf("%s%zu", i ? "," : "", inc[i]);
        printf("\n");
        return 0;
    }
    unsigned bits = argc > 1 ? atoi(argv[1]) : 2048, k = argc > 2 ? atoi(argv[2]) : 17;
    const std::string mode = argc > 3 ? argv[3] : "65537";
    // mode "digest" <e>: RSASignatureVerifier's byte composition + verify (src/lib.rs:183-248)
    unsigned long e = (mode == "var" || mode == "op") ? 0 : mode == "digest" ? strtoul(argc > 4 ? argv[4] : "65537", nullptr, 10) : strtoul(argv[3], nullptr, 10);
    std::vector<uint8_t> e_le;
    for (unsigned long v = e; v; v >>= 8) e_le.push_back((uint8_t)v);
    try {
        RegionCtx rc((1u << k) - BLINDING_ROWS);
        AssignedValue is_valid;
        if (mode == "var") is_valid = record_rsa_pkcs1v15_var(rc, bits, argc > 4 ? atoi(argv[4]) : 17);
        else if (mode == "op") is_valid = record_bigint_op(rc, (uint32_t)atoi(argv[4]), bits, argc > 5 ? atoi(argv[5]) : 5, nullptr);
        else if (mode == "digest") is_valid = record_rsa_verifier_from_digest(rc, bits, e_le);
        else is_valid = record_rsa_pkcs1v15(rc, bits, e_le);
        uint64_t hf = 0, hc = 0, hr = 0;
        for (uint32_t r = 0; r < rc.offset; r++) {
            for (int f = 0; f < NUM_FIXED; f++) {
                const U256& v = rc.constants[rc.fixed[r][f]];
                if (v.is_zero()) continue;
                uint64_t h = mix64(mix64(0xcbf29ce484222325ull, r), f);
                for (int i = 0; i < 4; i++) h = mix64(h, v.l[i]);
                hf += h;
            }
            const auto& t = rc.range_tags[r];
            uint64_t h = mix64(mix64(mix64(mix64(mix64(0xcbf29ce484222325ull, r), t[0]), t[1]), t[2]), t[3]);
            if (t[0] | t[2]) hr += h;
        }
        for (const auto& c : rc.copies) hc += mix64(mix64(mix64(mix64(0xcbf29ce484222325ull, c[0]), c[1]), c[2]), c[3]);
        uint32_t nlev = 0;
        for (uint32_t l : rc.level) nlev = l + 1 > nlev ? l + 1 : nlev;
        printf("rows=%u fixed=%llu ncopies=%zu copies=%llu range=%llu nodes=%zu levels=%u bigops=%zu is_valid_row=%u\n", rc.offset,
               (unsigned long long)hf, rc.copies.size(), (unsigned long long)hc, (unsigned long long)hr, rc.nodes.size(), nlev,
               rc.big_ops.size(), is_valid.row);
    } catch (const SynthError& err) {
        printf("error=%d %s\n", err.code, err.what());
        return 2;
    }
    return 0;
}
