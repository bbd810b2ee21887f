// fuzz_host.cpp -- mutation fuzzer for the product's HOST parsers (test tool).  No sanitizer runtime exists in this image, so
// every input (and the Zstandard output buffer) is placed so that it ENDS at an inaccessible guard page and starts right after
// one: any read or write past either end of a buffer is a segmentation fault.
// The host side parses untrusted bytes before anything reaches the GPU: the Draco header walk (draco_parse.cpp), the KTX2
// container parse (basis_parse.cpp) and the Zstandard decoder (zstd_inflate.cpp).  Every seed file given on the command line
// is mutated (bit flips, byte stores, truncations, splices of 32-bit extremes) `rounds` times; a sanitizer report or a
// crash fails the test, any status code is acceptable.   usage: fuzz_host <rounds> <seed> files...
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>
#include "../../universal-volumetric_b200/csrc/uvol_internal.h"

int uvol_draco_parse(const uint8_t *data, size_t len, DracoFrame &f, std::vector<uint32_t> &aux);
int uvol_ktx2_parse(const uint8_t *b, size_t len, uint32_t file_index, Ktx2File &f, std::vector<Ktx2Slice> &slices);
extern "C" int uvol_zstd_inflate(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *out_len);

static uint64_t rng_state;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

// [guard page][ ... n bytes ending at a page boundary][guard page]; `front` = true puts the bytes right after the first guard instead
struct Guarded { uint8_t *base = nullptr, *p = nullptr; size_t map = 0; };
static Guarded guarded(size_t n, bool front) {
    const size_t pg = (size_t)sysconf(_SC_PAGESIZE), body = (n + pg - 1) / pg * pg + (n == 0 ? pg : 0);
    Guarded g; g.map = body + 2 * pg;
    g.base = (uint8_t *)mmap(nullptr, g.map, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (g.base == (uint8_t *)MAP_FAILED) { perror("mmap"); exit(2); }
    mprotect(g.base, pg, PROT_NONE); mprotect(g.base + pg + body, pg, PROT_NONE);
    g.p = front ? g.base + pg : g.base + pg + body - n;
    return g;
}
static size_t g_zcap = 1 << 18;      // output capacity for Zstandard inputs: the EXACT content size of the current seed (as the product passes it)
static void run_one(const std::string &name, const uint8_t *p, size_t n, long *ok) {
    static long flip = 0;
    Guarded in = guarded(n, (flip++ & 7) == 7);          // mostly end-aligned (over-reads), sometimes start-aligned (under-reads)
    uint8_t *buf = in.p; memcpy(buf, p, n);
    int rc;
    if (name.size() > 4 && name.substr(name.size() - 4) == ".drc") { DracoFrame f; memset(&f, 0, sizeof f); std::vector<uint32_t> aux; rc = uvol_draco_parse(buf, n, f, aux); }
    else if (name.size() > 5 && name.substr(name.size() - 5) == ".ktx2") { Ktx2File f; memset(&f, 0, sizeof f); std::vector<Ktx2Slice> sl; rc = uvol_ktx2_parse(buf, n, 0, f, sl); }
    else { const size_t cap = g_zcap; Guarded out = guarded(cap, false); size_t got = 0; rc = uvol_zstd_inflate(buf, n, out.p, cap, &got); munmap(out.base, out.map); }
    if (rc == 0) ++*ok;
    munmap(in.base, in.map);
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int rounds = atoi(argv[1]); rng_state = 0x9E3779B97F4A7C15ull ^ (uint64_t)atoll(argv[2]);
    long total = 0, ok = 0;
    for (int a = 3; a < argc; a++) {
        FILE *fp = fopen(argv[a], "rb"); if (!fp) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
        std::vector<uint8_t> seed; uint8_t tmp[65536]; size_t k;
        while ((k = fread(tmp, 1, sizeof tmp, fp)) > 0) seed.insert(seed.end(), tmp, tmp + k);
        fclose(fp);
        const std::string name = argv[a];
        if (name.size() > 4 && name.substr(name.size() - 4) == ".zst") { std::vector<uint8_t> big(8u << 20); size_t got = 0; if (uvol_zstd_inflate(seed.data(), seed.size(), big.data(), big.size(), &got) == 0) g_zcap = got; else g_zcap = 1 << 18; }
        run_one(name, seed.data(), seed.size(), &ok); total++;
        for (int r = 0; r < rounds; r++) {
            std::vector<uint8_t> m = seed;
            const int kind = rnd() % 5, edits = 1 + rnd() % 4;
            // header-biased positions: most structure lives in the first few hundred bytes
            auto pos = [&]() { return m.empty() ? 0u : (rnd() % 3 ? rnd() % (uint32_t)(m.size() < 512 ? m.size() : 512) : rnd() % (uint32_t)m.size()); };
            for (int e = 0; e < edits && !m.empty(); e++) {
                if (kind == 0) m[pos()] ^= (uint8_t)(1u << (rnd() % 8));
                else if (kind == 1) m[pos()] = (uint8_t)rnd();
                else if (kind == 2) m.resize(rnd() % (m.size() + 1));
                else if (kind == 3) { const uint32_t v[6] = {0u, 0xFFFFFFFFu, 0x7FFFFFFFu, 0x80000000u, 0x00FFFFFFu, (uint32_t)m.size()}; const size_t at = pos(); for (int b = 0; b < 4 && at + b < m.size(); b++) m[at + b] = (uint8_t)(v[rnd() % 6] >> (8 * b)); }
                else { const size_t at = pos(), len = rnd() % 16; for (size_t b = 0; b < len && at + b < m.size(); b++) m[at + b] = (uint8_t)rnd(); }
            }
            run_one(name, m.data(), m.size(), &ok); total++;
        }
    }
    printf("fuzz_host: %ld inputs, %ld parsed ok, no access outside the buffers\n", total, ok);
    return 0;
}
