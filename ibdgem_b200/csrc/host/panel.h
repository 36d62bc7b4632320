// panel.h — host packer: genotype panel (IMPUTE .hap/.legend/.indv or VCF) + pileup -> the arrays
// the engine uploads (include/ibdgem_b200.h).  The panel is parsed ONCE, not once per target as
// the reference does (src/ibdgem.c:573-574, 771-772).  Also the list parsers of src/ibd-parse.c.
#pragma once

#include <cstdint>
#include <string>
#include <unordered_set>
#include <vector>

#include "pileup_store.h"

namespace ibdhost {

struct Sample {
    std::string name;
    int32_t ordinal;  // Sampl.idx / 2 (src/ibd-parse.c:29)
};

struct FreqTable {  // -A file (src/ibd-parse.c:311-358), position-sorted by contract
    std::vector<uint64_t> pos;
    std::vector<double> f;
    const double *fetch(uint64_t position) const;  // glibc bsearch probe order (src/ibd-parse.c:82-88)
};

struct PackedPanel {
    int64_t S = 0;   // panel lines (one per .hap/.legend line or VCF record)
    int32_t N = 0;   // individuals
    int64_t Wh = 0;  // 32-bit words per packed row, multiple of 4
    std::vector<std::string> names;
    std::vector<uint64_t> pos;
    std::vector<uint8_t> n_ref, n_alt, host_keep;
    std::vector<double> af_user;  // empty unless -A
    std::vector<uint32_t> bits;   // [S][Wh]
    // what the tab.txt row prints besides numbers (kept sites only)
    std::vector<uint32_t> dp;       // raw pileup coverage
    std::vector<uint32_t> chr_id;   // pileup chromosome name index
    std::vector<uint64_t> id_off;   // rsID text in `text`
    std::vector<uint32_t> id_len;
    std::vector<char> ref, alt;
    std::string text;
};

// --- lists -----------------------------------------------------------------------------------
int read_indv(const std::string &fn, std::vector<std::string> *names);                      // src/ibd-parse.c:4-42
// -S / -B files and the -s string: names not in the panel are warned about and dropped.
int read_sample_file(const std::string &fn, const std::vector<std::string> &names, bool background,
                     std::vector<Sample> *out);                                              // :176-214, 262-308
int read_sample_string(const std::string &s, const std::vector<std::string> &names, std::vector<Sample> *out);  // :217-259
int read_af(const std::string &fn, const char *chr, FreqTable *out);                        // :311-358
int read_positions(const std::string &fn, const char *chr, std::unordered_set<uint64_t> *out);  // :361-421
int find_sample(const std::vector<std::string> &names, const std::string &id);              // :45-52, -1 if absent

// --- packers ---------------------------------------------------------------------------------
struct PackOptions {
    const std::unordered_set<uint64_t> *positions = nullptr;  // -p
    const FreqTable *af = nullptr;                            // -A
    double min_qual = 0;                                      // -q (VCF)
};
int pack_impute(const std::string &hap_fn, const std::string &legend_fn, const std::vector<std::string> &names,
                const PileupStore &pu, const PackOptions &opt, PackedPanel *out);

// The pileup-independent half of a panel — what the .hap / .legend / .indv text (or the VCF) says — and
// its binary cache (one panel is scored against many pileups; re-parsing 10 GB of
// text per run is the reference's real wall-clock cost, src/ibdgem.c:573-574).
struct PanelText {
    int64_t S = 0;
    int32_t N = 0;
    int64_t Wh = 0;
    std::vector<std::string> names;
    std::vector<uint64_t> pos;    // 0 where the legend line did not parse (src/ibdgem.c:589)
    std::vector<uint8_t> state;   // 0 = line unparsable / skipped, 1 = parsed but not a SNP, 2 = SNP
    std::vector<uint64_t> id_off; // rsID of every parsed line, NUL-terminated, in `text`
    std::vector<uint32_t> id_len;
    std::vector<char> ref, alt;   // first character of REF / ALT
    std::vector<double> qual;     // VCF only: QUAL of the record (-q is applied when joining); empty for IMPUTE
    std::string text;
    std::vector<uint32_t> bits;   // [S][Wh]
};
int parse_impute(const std::string &hap_fn, const std::string &legend_fn, const std::vector<std::string> &names,
                 PanelText *out);
// Joins the panel text with a pileup and the option tables; takes the bits out of `pt`.
int join_pileup(PanelText *pt, const PileupStore &pu, const PackOptions &opt, PackedPanel *out);
int parse_vcf(const std::string &vcf_fn, PanelText *out);
// Cache file = the PanelText arrays behind a header that names size and mtime of the input files
// (.hap, .legend, .indv — or the VCF); a cache whose header does not match them is ignored and rewritten.
int save_panel_cache(const std::string &cache_fn, const std::vector<std::string> &inputs, const PanelText &pt);
bool load_panel_cache(const std::string &cache_fn, const std::vector<std::string> &inputs, PanelText *pt);
// pack_impute through the cache: *hit tells whether the text files were parsed (false) or not (true).
int pack_impute_cached(const std::string &hap_fn, const std::string &legend_fn, const std::string &indv_fn,
                       const std::string &cache_fn, const PileupStore &pu, const PackOptions &opt, PackedPanel *out,
                       bool *hit);
int pack_vcf_cached(const std::string &vcf_fn, const std::string &cache_fn, const PileupStore &pu, const PackOptions &opt,
                    PackedPanel *out, bool *hit);
int pack_vcf(const std::string &vcf_fn, const PileupStore &pu, const PackOptions &opt, PackedPanel *out);

}  // namespace ibdhost
