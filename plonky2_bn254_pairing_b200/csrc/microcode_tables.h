// Interface between the generated program tables (microcode_gen.cpp) and the runtime (bnp.cu).
#pragma once
#include <stdint.h>

struct BnpProgram {
    const char* name;
    const uint64_t* code;   // instruction words, END-terminated, padded for the prefetch
    uint32_t len;           // words including padding
    uint32_t n_slots;       // shared-memory Fq2 slots per thread
    uint32_t n_scratch;     // global scratch Fq2 slots per thread
    uint64_t macs;          // algorithmic 32x32 MACs per element (64/product + 72/reduction)
    uint32_t products;      // Fp wide products per element
    uint32_t reductions;    // Montgomery reductions per element
    uint32_t n_state;       // Fq2 values of per-element phase state (phase programs "name#K.i" only)
    uint64_t macs_executed; // MACs the component-split kernel issues per element (4-product Fq2 multiplication)
};

extern const uint32_t BNP_NCONST;
extern const uint32_t BNP_CONST_TABLE[][16];
extern const uint32_t BNP_NPROG;
extern const BnpProgram BNP_PROGRAMS[];
