// Host-side exerciser for the blinding stream of csrc/devutil.cuh (ChaCha20 block -> 512-bit reduction mod r).
// Prints the canonical value of a few (seed, nonce, proof, stream, row) cells; tests/test_transcript_host.py
// compares them with the oracle's independent restatement (oracle/plonk.py: blind_fe).
#include <cstdio>
#include <cstdlib>
#include "../../halo2-rsa_b200/csrc/devutil.cuh"
using namespace b2r;

static void pr(const fe_t& m) { fe_t c = Fr::from_mont(m); for (int i = 7; i >= 0; i--) printf("%08x", c.l[i]); printf("\n"); }

int main(int argc, char** argv) {
    // args: seed64 nonce proof stream row  (repeated)  |  "key" uses the 32-byte key 00 01 .. 1f
    for (int i = 1; i + 4 < argc; i += 5) {
        const uint64_t nonce = strtoull(argv[i + 1], nullptr, 0);
        BlindKey K;
        if (argv[i][0] == 'k') {
            uint8_t kb[32];
            for (int j = 0; j < 32; j++) kb[j] = (uint8_t)j;
            K = blind_key_from_bytes(kb, nonce);
        } else {
            K = blind_key_from_seed64(strtoull(argv[i], nullptr, 0), nonce);
        }
        pr(blind_value(K, (uint32_t)strtoul(argv[i + 2], nullptr, 0), (uint32_t)strtoul(argv[i + 3], nullptr, 0), (uint32_t)strtoul(argv[i + 4], nullptr, 0)));
    }
    return 0;
}
