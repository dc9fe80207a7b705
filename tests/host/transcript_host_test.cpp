// Host-side exerciser for csrc/transcript.hpp (Blake2b transcript of the prover).  Prints the
// challenges and the proof stream of a fixed script; tests/test_transcript_host.py replays the same
// script on the oracle's hashlib-based transcript (oracle/plonk.py) and compares.
#include <cstdio>
#include "../../halo2-rsa_b200/csrc/transcript.hpp"
using namespace b2r;

static fe_t fr_u64(uint64_t v) { fe_t c = Fr::zero(); c.l[0] = (uint32_t)v; c.l[1] = (uint32_t)(v >> 32); return Fr::to_mont(c); }
static fe_t fq_u64(uint64_t v) { fe_t c = Fq::zero(); c.l[0] = (uint32_t)v; c.l[1] = (uint32_t)(v >> 32); return Fq::to_mont(c); }
static void pr(const fe_t& m) { fe_t c = Fr::from_mont(m); for (int i = 7; i >= 0; i--) printf("%08x", c.l[i]); printf("\n"); }

int main() {
    Transcript t;
    t.common_scalar(fr_u64(123456789));
    t.write_point(fq_u64(1), fq_u64(2));
    fe_t c1 = t.squeeze();
    pr(c1);
    for (int i = 0; i < 40; i++) t.write_scalar(Fr::mul(c1, fr_u64(i + 7)));   // crosses several 128-byte blocks
    fe_t c2 = t.squeeze();
    pr(c2);
    pr(t.squeeze());
    t.write_point(Fq::zero(), Fq::zero());  // identity
    // a point with odd y: (1, q - 2)
    t.write_point(fq_u64(1), Fq::neg(fq_u64(2)));
    pr(t.squeeze());
    for (uint8_t b : t.out) printf("%02x", b);
    printf("\n");
    return 0;
}
