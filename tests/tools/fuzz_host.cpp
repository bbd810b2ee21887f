// fuzz_host.cpp -- mutation fuzzer for the product's HOST parsers (test tool).  No sanitizer runtime exists in this image, so
// every input (and the Zstandard output buffer) is placed so that it ENDS at an inaccessible guard page and starts right after
// one: any read or write past either end of a buffer is a segmentation fault.
// The host side parses untrusted bytes before anything reaches the GPU: the Draco header walk (draco_parse.cpp), the KTX2
// container parse (basis_parse.cpp) and the Zstandard decoder (zstd_inflate.cpp).  Every seed file given on the command line
// is mutated (bit flips, byte stores, truncations, splices of 32- and 64-bit extremes, over-long varints) `rounds` times and then
// run through DIRECTED cases (64-bit extremes at every KTX2 index field, the key/value length word, varint extremes at every
// position of a .drc's header and decoder sections).  The output descriptors (DracoFrame / Ktx2File) sit against a guard page as
// well, and every input runs under alarm(): a crash, a write past a descriptor or a hang fails the test, any status code is
// acceptable.   usage: fuzz_host <rounds> <seed> files...
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
int uvol_ktx2_split_levels(const uint8_t *b, size_t len, std::vector<std::vector<uint8_t>> &out);
#include "../../universal-volumetric_b200/csrc/corto_parse.h"
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
template <class T> struct GuardedObj {          // one T ending exactly at a guard page (writes past the descriptor fault)
    Guarded g; T *p;
    GuardedObj() : g(guarded(sizeof(T), false)), p((T *)g.p) { memset((void *)p, 0, sizeof(T)); }
    ~GuardedObj() { munmap(g.base, g.map); }
};
static size_t g_zcap = 1 << 18;      // output capacity for Zstandard inputs: the EXACT content size of the current seed (as the product passes it)
static void run_one(const std::string &name, const uint8_t *p, size_t n, long *ok) {
    static long flip = 0;
    Guarded in = guarded(n, (flip++ & 7) == 7);          // mostly end-aligned (over-reads), sometimes start-aligned (under-reads)
    uint8_t *buf = in.p; memcpy(buf, p, n);
    int rc;
    alarm(20);                                           // a parser that loops on a crafted length is a failure too (SIGALRM kills the run)
    if (name.size() > 4 && name.substr(name.size() - 4) == ".drc") { GuardedObj<DracoFrame> f; std::vector<uint32_t> aux; rc = uvol_draco_parse(buf, n, *f.p, aux); }
    else if (name.size() > 5 && name.substr(name.size() - 5) == ".ktx2") {
        {   // mip chains are taken apart first (basis_parse.cpp): the splitter sees the same untrusted bytes, and what it emits must be
            // single-level files whose descriptors stay inside THEM
            std::vector<std::vector<uint8_t>> lv; const int L = uvol_ktx2_split_levels(buf, n, lv);
            for (int k = 0; k < L; k++) {
                Ktx2File g; memset(&g, 0, sizeof g); std::vector<Ktx2Slice> s2;
                if (uvol_ktx2_parse(lv[k].data(), lv[k].size(), 0, g, s2) == 0) for (const Ktx2Slice &s : s2) if ((uint64_t)s.data_off + s.data_len > lv[k].size()) { fprintf(stderr, "fuzz_host: split level points outside itself\n"); abort(); }
            }
        }
        GuardedObj<Ktx2File> f; std::vector<Ktx2Slice> sl; rc = uvol_ktx2_parse(buf, n, 0, *f.p, sl);
        if (rc == 0) {          // what the device kernels will dereference must lie inside the file
            const Ktx2File &k = *f.p;
            bool in = true;
            for (const Ktx2Slice &s : sl) in = in && (uint64_t)s.data_off + s.data_len <= n;
            if (!k.is_uastc) in = in && (uint64_t)k.ep_off + k.ep_len <= n && (uint64_t)k.sel_off + k.sel_len <= n && (uint64_t)k.tab_off + k.tab_len <= n;
            else if (!k.zstd) in = in && (uint64_t)k.level_off + (uint64_t)k.layers * k.bx * k.by * 16 <= n;
            else in = in && (uint64_t)k.z_src_off + k.z_src_len <= n;
            if (!in) { fprintf(stderr, "fuzz_host: accepted KTX2 descriptor points outside the file\n"); abort(); }
        }
    }
    else if (name.size() > 4 && name.substr(name.size() - 4) == ".crt") {          // V1 Corto frame: header + section walk (corto_parse.h)
        GuardedObj<CortoFrame> f; std::vector<uint32_t> aux; rc = corto_parse(buf, n, *f.p, aux, 1ull << 24);
        if (rc == 0) {          // every section the kernels will read lies inside the file, and the counts are what the sections can hold
            const CortoFrame &c = *f.p; bool in = true;
            auto tun = [&](const TunBlock &t) { return (uint64_t)t.data_off + t.csize <= n && (t.nsym == 0xffffffffu || (uint64_t)t.probs_off + 2ull * t.nsym <= n); };
            auto bits = [&](const BitBlock &b) { return (uint64_t)b.data_off + 4ull * b.nwords <= n; };
            in = in && tun(c.clers) && bits(c.ibits) && c.nattr >= 1 && c.nattr <= CORTO_MAX_ATTRS && c.pos_attr >= 0 && c.pos_attr < c.nattr;
            for (int a = 0; a < c.nattr && in; a++) { in = in && bits(c.attr[a].bits) && c.attr[a].nlogs >= 1 && c.attr[a].nlogs <= 4; for (int k = 0; k < c.attr[a].nlogs && in; k++) in = in && tun(c.attr[a].logs[k]); }
            in = in && (uint64_t)c.nface <= (uint64_t)c.clers.size + 1 && c.groups_off + (uint64_t)c.ngroups <= aux.size();
            if (!in) { fprintf(stderr, "fuzz_host: accepted Corto descriptor points outside the file\n"); abort(); }
        }
    }
    else { const size_t cap = g_zcap; Guarded out = guarded(cap, false); size_t got = 0; rc = uvol_zstd_inflate(buf, n, out.p, cap, &got); munmap(out.base, out.map); }
    alarm(0);
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
            const int kind = rnd() % 7, edits = 1 + rnd() % 4;
            // header-biased positions: most structure lives in the first few hundred bytes
            auto pos = [&]() { return m.empty() ? 0u : (rnd() % 3 ? rnd() % (uint32_t)(m.size() < 512 ? m.size() : 512) : rnd() % (uint32_t)m.size()); };
            for (int e = 0; e < edits && !m.empty(); e++) {
                if (kind == 0) m[pos()] ^= (uint8_t)(1u << (rnd() % 8));
                else if (kind == 1) m[pos()] = (uint8_t)rnd();
                else if (kind == 2) m.resize(rnd() % (m.size() + 1));
                else if (kind == 3) { const uint32_t v[6] = {0u, 0xFFFFFFFFu, 0x7FFFFFFFu, 0x80000000u, 0x00FFFFFFu, (uint32_t)m.size()}; const size_t at = pos(); for (int b = 0; b < 4 && at + b < m.size(); b++) m[at + b] = (uint8_t)(v[rnd() % 6] >> (8 * b)); }
                else if (kind == 4) { const size_t at = pos(), len = rnd() % 16; for (size_t b = 0; b < len && at + b < m.size(); b++) m[at + b] = (uint8_t)rnd(); }
                else if (kind == 5) {          // 64-bit extremes (offset / length fields that wrap when summed)
                    const uint64_t v[7] = {~0ull, ~0ull - (1ull << 30) + 1, 1ull << 63, 1ull << 32, (1ull << 32) - 1, (uint64_t)m.size(), ~0ull - (uint64_t)m.size() + 9};
                    const uint64_t x = v[rnd() % 7]; const size_t at = pos() & ~(size_t)3; for (int b = 0; b < 8 && at + b < m.size(); b++) m[at + b] = (uint8_t)(x >> (8 * b));
                } else {                       // over-long varints (INT_MAX, UINT_MAX, 2^63) spliced over a count field
                    static const uint8_t vi[3][10] = {{0xFF, 0xFF, 0xFF, 0xFF, 0x07}, {0xFF, 0xFF, 0xFF, 0xFF, 0x0F}, {0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x01}};
                    static const int vl[3] = {5, 5, 10};
                    const int w = rnd() % 3; const size_t at = pos(); m.erase(m.begin() + at); m.insert(m.begin() + at, vi[w], vi[w] + vl[w]);
                }
            }
            run_one(name, m.data(), m.size(), &ok); total++;
        }
    }
    // ---- directed cases
    for (int a = 3; a < argc; a++) {
        FILE *fp = fopen(argv[a], "rb"); if (!fp) return 2;
        std::vector<uint8_t> seed; uint8_t tmp[65536]; size_t k;
        while ((k = fread(tmp, 1, sizeof tmp, fp)) > 0) seed.insert(seed.end(), tmp, tmp + k);
        fclose(fp);
        const std::string name = argv[a];
        if (name.size() > 5 && name.substr(name.size() - 5) == ".ktx2" && seed.size() >= 104) {
            const uint64_t L = seed.size();
            const uint64_t v64[10] = {~0ull, ~0ull - (1ull << 30) + 1, (1ull << 30) + 5638, 1ull << 63, 1ull << 32, L, L - 1, ~0ull - L + 9, 0ull, 1ull << 31};
            const int off64[5] = {64, 72, 80, 88, 96}, off32[4] = {48, 52, 56, 60};
            for (int i = 0; i < 5; i++) for (int j = 0; j < 10; j++) for (int i2 = -1; i2 < 5; i2++) for (int j2 = 0; j2 < (i2 < 0 ? 1 : 10); j2++) {
                std::vector<uint8_t> m = seed; memcpy(&m[off64[i]], &v64[j], 8); if (i2 >= 0) memcpy(&m[off64[i2]], &v64[j2], 8);
                run_one(name, m.data(), m.size(), &ok); total++;
            }
            uint32_t nlev; memcpy(&nlev, &seed[40], 4);
            if (nlev > 1 && nlev < 16 && seed.size() >= 80 + 24ull * nlev) {          // a mip chain: every entry of the level index, and the level / layer counts
                for (uint32_t e = 0; e < 3 * nlev; e++) for (int j = 0; j < 10; j++) { std::vector<uint8_t> m = seed; memcpy(&m[80 + 8 * e], &v64[j], 8); run_one(name, m.data(), m.size(), &ok); total++; }
                const uint32_t cnt[6] = {0u, 2u, 15u, 16u, 0x7FFFFFFFu, 0xFFFFFFFFu};
                for (int f2 = 0; f2 < 2; f2++) for (int j = 0; j < 6; j++) { std::vector<uint8_t> m = seed; memcpy(&m[f2 ? 32 : 40], &cnt[j], 4); run_one(name, m.data(), m.size(), &ok); total++; }
            }
            const uint32_t v32[6] = {0xFFFFFFFFu, 0xFFFFFFFCu, 0x7FFFFFFFu, 0x80000000u, (uint32_t)L, 0u};
            for (int i = 0; i < 4; i++) for (int j = 0; j < 6; j++) { std::vector<uint8_t> m = seed; memcpy(&m[off32[i]], &v32[j], 4); run_one(name, m.data(), m.size(), &ok); total++; }
            uint32_t kvd; memcpy(&kvd, &seed[56], 4);
            if ((uint64_t)kvd + 4 <= L) for (int j = 0; j < 6; j++) { std::vector<uint8_t> m = seed; memcpy(&m[kvd], &v32[j], 4); run_one(name, m.data(), m.size(), &ok); total++; }   // keyAndValueByteLength
        }
        if (name.size() > 4 && name.substr(name.size() - 4) == ".crt") {          // every 32-bit field of the header / section heads with extreme values
            const uint32_t v32[7] = {0xFFFFFFFFu, 0x7FFFFFFFu, 0x80000000u, (uint32_t)seed.size(), 1u << 27, 1u << 26, 0u};
            const size_t lim = seed.size() < 4096 ? seed.size() : 4096;
            for (size_t at = 0; at + 4 <= lim; at++) for (int j = 0; j < 7; j++) { std::vector<uint8_t> m = seed; memcpy(&m[at], &v32[j], 4); run_one(name, m.data(), m.size(), &ok); total++; }
        }
        if (name.size() > 4 && name.substr(name.size() - 4) == ".drc") {
            static const uint8_t vi[3][5] = {{0xFF, 0xFF, 0xFF, 0xFF, 0x07}, {0xFF, 0xFF, 0xFF, 0xFF, 0x0F}, {0xFE, 0xFF, 0xFF, 0xFF, 0x07}};
            const size_t lim = seed.size() < 16384 ? seed.size() : 16384;
            for (size_t at = 11; at < lim; at++) {
                if (seed[at] > 16) continue;                         // count-like bytes only
                for (int w = 0; w < 3; w++) {
                    std::vector<uint8_t> m = seed; m.erase(m.begin() + at); m.insert(m.begin() + at, vi[w], vi[w] + 5);
                    static const uint8_t rec[5] = {0, 9, 3, 0, 0};      // followed by plausible attribute records, so the loop keeps writing
                    for (int r = 0; r < 40 && at + 5 + 5 * r + 5 <= m.size(); r++) memcpy(&m[at + 5 + 5 * r], rec, 5);
                    run_one(name, m.data(), m.size(), &ok); total++;
                }
            }
        }
    }
    printf("fuzz_host: %ld inputs, %ld parsed ok, no access outside the buffers\n", total, ok);
    return 0;
}
