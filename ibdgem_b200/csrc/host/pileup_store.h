// pileup_store.h — in-memory store of one samtools mpileup file (replaces Pul / Pu_chr,
// init_Pu_chr, line2pul, fetch_Pul and count_base_from_pul, src/pileup.c:206-415, 442-559).
// Host-side packer code: it feeds the integer inputs of the engine (n_ref, n_alt, DP).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace ibdhost {

struct PileupStore {
    std::vector<uint32_t> pos;      // unsigned int in the reference (src/pileup.h:20)
    std::vector<uint32_t> cov;      // raw coverage, < 128
    std::vector<uint32_t> chr_id;   // index into chr_names
    std::vector<uint64_t> base_off; // offset of the line's bases in `bases`
    std::string bases;              // cov characters per line: A C G T N *
    std::vector<std::string> chr_names;

    size_t size() const { return pos.size(); }
    // bsearch with glibc's probe order, so duplicate positions resolve as in the reference
    // (src/pileup.c:472-485); -1 if absent.
    int64_t fetch(uint64_t position) const;
    // occurrences of `base` among the line's bases (src/pileup.c:442-450)
    unsigned count_base(int64_t line, char base) const;
};

// Reads and filters the pileup as init_Pu_chr does (chr == nullptr: no chromosome filter).
// Returns 0, or 1 after printing the reference's diagnostics (nothing parsed, unsorted, cannot open).
int load_pileup(const std::string &fn, const char *chr, PileupStore *out);

}  // namespace ibdhost
