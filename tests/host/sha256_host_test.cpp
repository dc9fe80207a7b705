// Host build of the device SHA-256 (halo2-rsa_b200/csrc/sha256.cuh): reads hex messages (one per line, "-" = empty) from
// stdin and prints "<digest hex> <limb0> <limb1> <limb2> <limb3>" per message.  Test scaffolding, not a CPU fallback.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../halo2-rsa_b200/csrc/sha256.cuh"

int main() {
    static char line[1 << 21];
    while (fgets(line, sizeof line, stdin)) {
        size_t n = strlen(line);
        while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        std::vector<uint8_t> m;
        if (strcmp(line, "-") != 0)
            for (size_t i = 0; i + 1 < n; i += 2) {
                unsigned v;
                sscanf(line + i, "%2x", &v);
                m.push_back((uint8_t)v);
            }
        uint32_t st[8];
        b2r::sha256_message(m.data(), m.size(), st);
        uint8_t d[32];
        uint64_t l[4];
        b2r::sha256_state_to_digest(st, d);
        b2r::sha256_state_to_limbs(st, l);
        for (int i = 0; i < 32; i++) printf("%02x", d[i]);
        printf(" %llx %llx %llx %llx\n", (unsigned long long)l[0], (unsigned long long)l[1], (unsigned long long)l[2], (unsigned long long)l[3]);
    }
    return 0;
}
