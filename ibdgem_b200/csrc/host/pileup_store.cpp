#include "pileup_store.h"

#include <cctype>
#include <cstdio>
#include <cstring>

#include "textio.h"

namespace ibdhost {

namespace {
constexpr unsigned kMaxCov = 128;      // src/pileup.h:12
constexpr size_t kFieldWidth = 10240;  // src/pileup.h:11

// One mpileup line -> (chr, pos, cov, bases).  Status as line2pul: 0 ok, 1 dropped silently
// (or with the reference's message), 2 unparsable start of line.
int parse_line(char *line, char *chr, unsigned *pos, unsigned *cov, char *bases, char *f_bases, char *f_bq, char *f_mq) {
    char ref;
    if (sscanf(line, "%255s\t%u\t%c\t%u\t", chr, pos, &ref, cov) != 4) return 2;
    if (*cov >= kMaxCov) return 1;  // src/pileup.c:223
    if (sscanf(line, "%255s\t%u\t%c\t%u\t%10239s\t%10239s\t%10239s", chr, pos, &ref, cov, f_bases, f_bq, f_mq) != 7) return 1;
    if (*cov == 0) return 0;  // special line with no real data
    const size_t n = strlen(f_bases);
    size_t i = 0;
    unsigned nb = 0;
    while (i < n) {
        const char c = f_bases[i];
        char b = 0;
        switch (c) {
            case '.': case ',': b = ref; break;  // the pileup's own reference column
            case 'A': case 'a': b = 'A'; break;
            case 'C': case 'c': b = 'C'; break;
            case 'G': case 'g': b = 'G'; break;
            case 'T': case 't': b = 'T'; break;
            case 'N': case 'n': b = 'N'; break;
            case '*': b = '*'; break;  // deletion marker, counted like a base
            case '-': case '+': {      // indel: skipped with its sequence
                i++;
                size_t len = 0;
                while (i < n && isdigit((unsigned char)f_bases[i])) len = len * 10 + (size_t)(f_bases[i++] - '0');
                i += len;
                continue;
            }
            case '$': i++; continue;
            case '^': i += 2; continue;  // skips the mapping-quality character too
            default:
                fprintf(stderr, "Cannot parse %c in reads field\n", c);
                return 1;
        }
        if (nb < kMaxCov) bases[nb] = b;
        nb++;
        i++;
    }
    if (nb != *cov) {
        fprintf(stderr, "Incorrect number of bases read in: %s\n", line);
        return 1;
    }
    if (strlen(f_bq) != nb && strlen(f_mq) != nb) {
        fprintf(stderr, "Incorrect number of base or map quals in: %s\n", line);
        return 1;
    }
    return 0;
}
}  // namespace

int64_t PileupStore::fetch(uint64_t position) const {
    size_t lo = 0, hi = pos.size();
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (position < pos[mid]) hi = mid;
        else if (position > pos[mid]) lo = mid + 1;
        else return (int64_t)mid;
    }
    return -1;
}

unsigned PileupStore::count_base(int64_t line, char base) const {
    unsigned c = 0;
    const char *b = bases.data() + base_off[(size_t)line];
    for (uint32_t i = 0; i < cov[(size_t)line]; i++) c += b[i] == base;
    return c;
}

int load_pileup(const std::string &fn, const char *chr, PileupStore *out) {
    LineReader lr;
    if (!lr.open(fn)) return 1;
    std::vector<char> buf, fb(kFieldWidth), fq(kFieldWidth), fm(kFieldWidth);
    char chrbuf[256], bases[kMaxCov];
    const char *line;
    size_t len;
    std::string last_chr;
    uint32_t last_id = 0;
    while (lr.next(&line, &len)) {
        buf.assign(line, line + len);
        buf.push_back('\0');
        unsigned pos = 0, cov = 0;
        const int st = parse_line(buf.data(), chrbuf, &pos, &cov, bases, fb.data(), fq.data(), fm.data());
        if (st == 2) {
            fprintf(stderr, "Problem parsing %s\n", buf.data());
            continue;
        }
        if (st) continue;
        if (chr && strcmp(chrbuf, chr) != 0) continue;
        if (out->chr_names.empty() || last_chr != chrbuf) {
            last_chr = chrbuf;
            last_id = (uint32_t)out->chr_names.size();
            for (uint32_t k = 0; k < out->chr_names.size(); k++)
                if (out->chr_names[k] == last_chr) last_id = k;
            if (last_id == out->chr_names.size()) out->chr_names.push_back(last_chr);
        }
        out->pos.push_back(pos);
        out->cov.push_back(cov);
        out->chr_id.push_back(last_id);
        out->base_off.push_back((uint64_t)out->bases.size());
        out->bases.append(bases, cov);
    }
    if (out->pos.empty()) {
        fprintf(stderr, "[::] ERROR in init_Pu_chr(): Cannot parse mpileup lines from %s.\n", fn.c_str());
        return 1;
    }
    for (size_t i = 0; i + 1 < out->pos.size(); i++)
        if (out->pos[i] > out->pos[i + 1]) {
            fprintf(stderr, "mpileup lines not sorted!\n");
            return 1;
        }
    return 0;
}

}  // namespace ibdhost
