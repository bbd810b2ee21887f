// corto_emu.cpp -- HOST EMULATION of the V1 connectivity walk (csrc/corto_core.h) for logic checks without a GPU.  Test tool only;
// never part of libuvol_b200.so, never a fallback.  Input = the CLERS symbols and split bit stream of a .crt (as the reference's
// own IndexAttribute::decode leaves them, oracle/_ref), output = faces + parallelogram contexts to compare with the reference's.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../universal-volumetric_b200/csrc/corto_core.h"

// ring_size: how many recent front records the walk keeps in its ring (the kernel: 1024; tests also run tiny rings so that every
// ring / global path is taken)
static int g_ring_size = 1024;
extern "C" void corto_emu_set_ring(int n) { g_ring_size = n; }
extern "C" int corto_emu_walk(const uint8_t *clers, uint32_t nclers, const uint32_t *words, uint32_t nwords, const uint32_t *group_end, uint32_t ngroups,
                              int nvert, int nface, uint32_t *faces, int *pred4) {
    std::vector<unsigned long long> cl((nclers + 15) / 8 + 1, 0); memcpy(cl.data(), clers, nclers);
    std::vector<uint32_t> w(nwords + 2, 0); if (nwords) memcpy(w.data(), words, (size_t)nwords * 4);
    const int cap = 3 * nface + 8;
    std::vector<CortoEdge> front(cap); std::vector<uint32_t> third(cap); std::vector<int> queue(cap), delayed(cap);
    std::vector<CortoEdge> ring((size_t)g_ring_size);
    CortoWalkMem m; m.ring = ring.data(); m.ring_size = g_ring_size; m.progress = nullptr;
    m.clers = (const uint8_t *)cl.data(); m.nclers = nclers; m.bits = CortoBits{w.data(), 0, (uint64_t)nwords * 32};
    m.group_end = group_end; m.ngroups = ngroups; m.front = front.data(); m.third = third.data(); m.front_cap = cap;
    m.queue = queue.data(); m.delayed = delayed.data(); m.order_cap = cap; m.faces = faces; m.pred = pred4; m.nvert = nvert; m.nface = nface;
    return corto_walk(m);
}
