#include "panel.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "textio.h"

namespace ibdhost {

namespace {

std::string zline(const char *line, size_t len) { return std::string(line, len); }

// the reference's "trim off new line char": the LAST character of the line goes, whatever it is
std::string drop_last(const char *line, size_t len) { return len ? std::string(line, len - 1) : std::string(); }

bool is_snp(const char *ref, const char *alt) {  // src/ibdgem.c:113-119
    return strlen(ref) == 1 && strchr("ACGT", ref[0]) && strlen(alt) == 1 && strchr("ACGT", alt[0]);
}

void init_panel(PackedPanel *p, int32_t n_indiv) {
    p->N = n_indiv;
    const int64_t words = (2 * (int64_t)n_indiv + 31) / 32;
    p->Wh = (words + 3) / 4 * 4;  // 16-byte rows for 128-bit loads on the device
}

void push_site(PackedPanel *p) {
    p->pos.push_back(0);
    p->n_ref.push_back(0);
    p->n_alt.push_back(0);
    p->host_keep.push_back(0);
    p->dp.push_back(0);
    p->chr_id.push_back(0);
    p->id_off.push_back(0);
    p->id_len.push_back(0);
    p->ref.push_back('.');
    p->alt.push_back('.');
    p->bits.resize(p->bits.size() + (size_t)p->Wh, 0u);
    p->S++;
}

// Everything about a kept line that comes from the pileup and the option tables.
void fill_kept(PackedPanel *p, size_t s, const PileupStore &pu, int64_t pul, const char *id, char ref, char alt) {
    p->host_keep[s] = 1;
    p->n_ref[s] = (uint8_t)pu.count_base(pul, ref);
    p->n_alt[s] = (uint8_t)pu.count_base(pul, alt);
    p->dp[s] = pu.cov[(size_t)pul];
    p->chr_id[s] = pu.chr_id[(size_t)pul];
    p->id_off[s] = p->text.size();
    p->id_len[s] = (uint32_t)strlen(id);
    p->text.append(id);
    p->ref[s] = ref;
    p->alt[s] = alt;
}

}  // namespace

const double *FreqTable::fetch(uint64_t position) const {
    size_t lo = 0, hi = pos.size();
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (position < pos[mid]) hi = mid;
        else if (position > pos[mid]) lo = mid + 1;
        else return &f[mid];
    }
    return nullptr;
}

int find_sample(const std::vector<std::string> &names, const std::string &id) {
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == id) return (int)i;
    return -1;
}

int read_indv(const std::string &fn, std::vector<std::string> *names) {
    LineReader lr;
    if (!lr.open(fn)) return 1;
    const char *line;
    size_t len;
    while (lr.next(&line, &len)) names->push_back(drop_last(line, len));
    if (names->empty()) {
        fprintf(stderr, "[::] ERROR: No samples found in .indv file.\n");
        return 1;
    }
    return 0;
}

int read_sample_file(const std::string &fn, const std::vector<std::string> &names, bool background,
                     std::vector<Sample> *out) {
    LineReader lr;
    if (!lr.open(fn)) return 1;
    const char *line;
    size_t len;
    while (lr.next(&line, &len)) {
        const std::string name = drop_last(line, len);
        const int k = find_sample(names, name);
        if (k < 0) {
            fprintf(stderr, background ? "Reference sample %s not found in input panel.\n" : "Sample %s not found in reference panel.\n",
                    name.c_str());
            continue;
        }
        out->push_back({name, k});
    }
    if (out->empty()) {
        fprintf(stderr, "[::] ERROR in %s(): No matching samples found in %s.\n", background ? "read_rf" : "read_sf", fn.c_str());
        return 1;
    }
    return 0;
}

int read_sample_string(const std::string &s, const std::vector<std::string> &names, std::vector<Sample> *out) {
    size_t i = 0;
    while (i < s.size()) {  // strtok(",") semantics: empty tokens vanish
        while (i < s.size() && s[i] == ',') i++;
        size_t j = i;
        while (j < s.size() && s[j] != ',') j++;
        if (j > i) {
            const std::string name = s.substr(i, j - i);
            const int k = find_sample(names, name);
            if (k < 0) fprintf(stderr, "Sample %s not found in reference panel.\n", name.c_str());
            else out->push_back({name, k});
        }
        i = j;
    }
    if (out->empty()) {
        fprintf(stderr, "[::] ERROR in read_scmd(): No matching samples found.\n");
        return 1;
    }
    return 0;
}

int read_af(const std::string &fn, const char *chr, FreqTable *out) {
    LineReader lr;
    if (!lr.open(fn)) return 1;
    const char *line;
    size_t len;
    char c[129];
    while (lr.next(&line, &len)) {
        const std::string z = zline(line, len);
        unsigned long pos;
        double f;
        if (sscanf(z.c_str(), "%128s %lu %lf", c, &pos, &f) == 3 && (!chr || strcmp(c, chr) == 0)) {
            out->pos.push_back(pos);
            out->f.push_back(f);
        }
    }
    if (out->pos.empty()) {
        fprintf(stderr, "[::] ERROR in read_af(): Cannot parse lines from %s.\n", fn.c_str());
        return 1;
    }
    return 0;
}

int read_positions(const std::string &fn, const char *chr, std::unordered_set<uint64_t> *out) {
    LineReader lr;
    if (!lr.open(fn)) return 1;
    const char *line;
    size_t len;
    char c[129];
    size_t n = 0;
    while (lr.next(&line, &len)) {
        const std::string z = zline(line, len);
        unsigned long pos;
        // BED (third column, 1-based end) first, then CHROM POS
        if (sscanf(z.c_str(), "%128s %*d %lu", c, &pos) == 2 || sscanf(z.c_str(), "%128s %lu", c, &pos) == 2) {
            if (!chr || strcmp(c, chr) == 0) {
                out->insert(pos);
                n++;
            }
        }
    }
    if (n == 0) {
        fprintf(stderr, "[::] ERROR in read_pos(): Cannot parse lines from %s.\n", fn.c_str());
        return 1;
    }
    return 0;
}

int pack_impute(const std::string &hap_fn, const std::string &legend_fn, const std::vector<std::string> &names,
                const PileupStore &pu, const PackOptions &opt, PackedPanel *out) {
    LineReader hap, leg;
    if (!hap.open(hap_fn) || !leg.open(legend_fn)) {
        fprintf(stderr, "[::] ERROR parsing hap/legend/indv data; make sure inputs are valid.\n");
        return 1;
    }
    out->names = names;
    init_panel(out, (int32_t)names.size());
    const size_t H = 2 * names.size();
    const char *hl, *ll;
    size_t hn, ln;
    leg.next(&ll, &ln);  // header (src/ibdgem.c:554)
    std::string lz;
    char id[129], ref[129], alt[129];
    while (hap.next(&hl, &hn) && leg.next(&ll, &ln)) {
        const size_t s = (size_t)out->S;
        push_site(out);
        if (hn < 2 * H - 1) {
            fprintf(stderr, "[::] ERROR: .hap line %zu has %zu characters, %zu haplotypes need %zu.\n", s + 1, hn, H, 2 * H - 1);
            return 1;
        }
        uint32_t *row = out->bits.data() + s * (size_t)out->Wh;
        size_t h = 0;
        // fast path: eight text bytes "a b c d " carry four alleles; check the pattern, pick bit 0 of
        // the four digits and gather them with one multiply
        for (; h + 4 <= H && 2 * h + 8 <= hn; h += 4) {
            uint64_t x;
            memcpy(&x, hl + 2 * h, 8);
            if ((x & 0xFFFEFFFEFFFEFFFEull) != 0x2030203020302030ull) break;  // not "[01] [01] [01] [01] "
            const uint64_t nib = (((x & 0x0001000100010001ull) * 0x0001000200040008ull) >> 48) & 0xFu;
            row[h >> 5] |= (uint32_t)nib << (h & 31);
        }
        for (; h < H; h++) {
            const char c = hl[2 * h];
            if (c == '1') row[h >> 5] |= 1u << (h & 31);
            else if (c != '0') {
                fprintf(stderr, "[::] ERROR: .hap line %zu: allele '%c' of haplotype %zu is not 0 or 1.\n", s + 1, c, h);
                return 1;
            }
        }
        lz.assign(ll, ln);
        unsigned long pos;
        if (sscanf(lz.c_str(), "%128s %lu %128s %128s", id, &pos, ref, alt) != 4) continue;  // src/ibdgem.c:589
        out->pos[s] = pos;
        if (!is_snp(ref, alt)) continue;
        const int64_t pul = pu.fetch(pos);
        if (pul < 0) continue;
        if (opt.positions && !opt.positions->count(pos)) continue;
        fill_kept(out, s, pu, pul, id, ref[0], alt[0]);
    }
    if (opt.af) {
        out->af_user.assign((size_t)out->S, NAN);
        for (int64_t s = 0; s < out->S; s++)
            if (out->host_keep[(size_t)s])
                if (const double *f = opt.af->fetch(out->pos[(size_t)s])) out->af_user[(size_t)s] = *f;
    }
    return 0;
}

int pack_vcf(const std::string &vcf_fn, const PileupStore &pu, const PackOptions &opt, PackedPanel *out) {
    LineReader vcf;
    if (!vcf.open(vcf_fn)) return 1;
    const char *line;
    size_t len;
    bool have = vcf.next(&line, &len);
    while (have && len >= 2 && line[0] == '#' && line[1] == '#') have = vcf.next(&line, &len);
    static const char kHead[] = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t";
    const size_t hl = sizeof(kHead) - 1;
    if (!have || len <= hl || memcmp(line, kHead, hl) != 0 || line[hl] == '\n') {
        fprintf(stderr, "[::] ERROR parsing VCF header.\n");
        return 1;
    }
    {  // sample names, split on tabs like strtok (src/ibd-parse.c:113-147)
        size_t e = len;
        while (e > hl && line[e - 1] == '\n') e--;
        size_t i = hl;
        while (i < e) {
            while (i < e && line[i] == '\t') i++;
            size_t j = i;
            while (j < e && line[j] != '\t') j++;
            if (j > i) out->names.emplace_back(line + i, j - i);
            i = j;
        }
        if (out->names.empty()) {
            fprintf(stderr, "[::] ERROR: No samples found.\n");
            return 1;
        }
    }
    init_panel(out, (int32_t)out->names.size());
    const size_t N = out->names.size();
    std::string id, ref, alt, qual;
    while (vcf.next(&line, &len)) {
        const size_t s = (size_t)out->S;
        push_site(out);
        size_t e = len;
        while (e > 0 && line[e - 1] == '\n') e--;
        // nine tab-separated leading fields, then the genotype columns (src/ibdgem.c:272-273)
        size_t fb[10], fe[10];
        size_t i = 0;
        int nf = 0;
        while (nf < 9 && i <= e) {
            size_t j = i;
            while (j < e && line[j] != '\t') j++;
            fb[nf] = i;
            fe[nf] = j;
            nf++;
            if (j >= e) { i = e + 1; break; }
            i = j + 1;
        }
        if (nf < 9 || i > e) continue;  // fewer than 10 columns: "skipped"
        char *endp = nullptr;
        const std::string posz(line + fb[1], fe[1] - fb[1]);
        const unsigned long pos = strtoul(posz.c_str(), &endp, 10);
        if (endp == posz.c_str()) continue;
        out->pos[s] = pos;
        id.assign(line + fb[2], fe[2] - fb[2]);
        ref.assign(line + fb[3], fe[3] - fb[3]);
        alt.assign(line + fb[4], fe[4] - fb[4]);
        qual.assign(line + fb[5], fe[5] - fb[5]);
        if (id.empty() || ref.empty() || alt.empty() || qual.empty() || fe[6] == fb[6] || fe[7] == fb[7]) continue;
        if (alt.find(',') != std::string::npos) continue;  // is_biallelic, src/ibdgem.c:161-167
        // genotypes of EVERY sample must look like [01][/|][01].* (src/ibd-parse.c:150-173)
        uint32_t *row = out->bits.data() + s * (size_t)out->Wh;
        size_t k = 0;
        bool ok = true;
        while (k < N) {
            while (i < e && line[i] == '\t') i++;
            size_t j = i;
            while (j < e && line[j] != '\t') j++;
            if (j - i < 3 || (line[i] != '0' && line[i] != '1') || (line[i + 1] != '/' && line[i + 1] != '|') ||
                (line[i + 2] != '0' && line[i + 2] != '1')) {
                ok = false;
                break;
            }
            if (line[i] == '1') row[(2 * k) >> 5] |= 1u << ((2 * k) & 31);
            if (line[i + 2] == '1') row[(2 * k + 1) >> 5] |= 1u << ((2 * k + 1) & 31);
            k++;
            i = j;
        }
        if (!ok) {
            fprintf(stderr, "Failed to parse genotype fields at %lu. Skipping to next site.\n", pos);
            for (int64_t w = 0; w < out->Wh; w++) row[w] = 0;
            continue;
        }
        if (!is_snp(ref.c_str(), alt.c_str())) continue;
        if (atof(qual.c_str()) < opt.min_qual) continue;
        const int64_t pul = pu.fetch(pos);
        if (pul < 0) continue;
        if (opt.positions && !opt.positions->count(pos)) continue;
        fill_kept(out, s, pu, pul, id.c_str(), ref[0], alt[0]);
    }
    if (opt.af) {
        out->af_user.assign((size_t)out->S, NAN);
        for (int64_t s = 0; s < out->S; s++)
            if (out->host_keep[(size_t)s])
                if (const double *f = opt.af->fetch(out->pos[(size_t)s])) out->af_user[(size_t)s] = *f;
    }
    return 0;
}

}  // namespace ibdhost
