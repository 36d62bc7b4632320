#include "pileup_store.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "textio.h"

namespace ibdhost {

namespace {
constexpr unsigned kMaxCov = 128;      // src/pileup.h:12
constexpr size_t kFieldWidth = 10240;  // src/pileup.h:11

// The seven fields of a line in the shape samtools writes: non-empty runs of non-blank characters
// separated by single tabs, position and coverage of at most nine digits, a one-character reference
// column.  On exactly these lines the two sscanf calls of line2pul (src/pileup.c:214-232) read the
// same values, so they can be skipped; anything else goes through sscanf itself.
struct Fields {
    const char *bases, *bq, *mq;
    size_t n_bases, n_bq, n_mq;
};
inline bool blank(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
bool split_fast(const char *p, const char *end, char *chr, unsigned *pos, char *ref, unsigned *cov, Fields *f) {
    auto run = [&](const char **b, size_t *n, size_t max) {
        const char *q = p;
        while (q < end && !blank(*q)) q++;
        *b = p;
        *n = (size_t)(q - p);
        p = q;
        return *n >= 1 && *n <= max;
    };
    auto tab = [&] { return p < end && *p++ == '\t'; };
    auto number = [&](unsigned *v) {
        const char *q = p;
        unsigned x = 0;
        while (q < end && *q >= '0' && *q <= '9' && q - p < 10) x = x * 10 + (unsigned)(*q++ - '0');
        const size_t nd = (size_t)(q - p);
        p = q;
        *v = x;
        return nd >= 1 && nd <= 9;
    };
    const char *c;
    size_t nc;
    if (memchr(p, 0, (size_t)(end - p))) return false;  // sscanf would stop at the NUL
    if (!run(&c, &nc, 255) || !tab() || !number(pos) || !tab()) return false;
    if (p >= end || blank(*p)) return false;
    *ref = *p++;
    if (!tab() || !number(cov) || !tab()) return false;
    if (!run(&f->bases, &f->n_bases, kFieldWidth - 1) || !tab() || !run(&f->bq, &f->n_bq, kFieldWidth - 1) || !tab() ||
        !run(&f->mq, &f->n_mq, kFieldWidth - 1))
        return false;
    if (p < end && !blank(*p)) return false;
    memcpy(chr, c, nc);
    chr[nc] = 0;
    return true;
}

// One mpileup line -> (chr, pos, cov, bases).  Status as line2pul: 0 ok, 1 dropped silently
// (or with the reference's message), 2 unparsable start of line.
int parse_line(char *line, size_t len, bool fast, char *chr, unsigned *pos, unsigned *cov, char *bases, char *f_bases, char *f_bq,
               char *f_mq) {
    char ref;
    Fields fl;
    if (fast && split_fast(line, line + len, chr, pos, &ref, cov, &fl)) {
        if (*cov >= kMaxCov) return 1;  // src/pileup.c:223
    } else {
        if (sscanf(line, "%255s\t%u\t%c\t%u\t", chr, pos, &ref, cov) != 4) return 2;
        if (*cov >= kMaxCov) return 1;  // src/pileup.c:223
        if (sscanf(line, "%255s\t%u\t%c\t%u\t%10239s\t%10239s\t%10239s", chr, pos, &ref, cov, f_bases, f_bq, f_mq) != 7) return 1;
        fl.bases = f_bases; fl.n_bases = strlen(f_bases);
        fl.bq = f_bq; fl.n_bq = strlen(f_bq);
        fl.mq = f_mq; fl.n_mq = strlen(f_mq);
    }
    if (*cov == 0) return 0;  // special line with no real data
    const size_t n = fl.n_bases;
    const char *rb = fl.bases;
    size_t i = 0;
    unsigned nb = 0;
    while (i < n) {
        const char c = rb[i];
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
                while (i < n && isdigit((unsigned char)rb[i])) len = len * 10 + (size_t)(rb[i++] - '0');
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
    if (fl.n_bq != nb && fl.n_mq != nb) {
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
    const bool fast = getenv("IBDGEM_PILEUP_NO_FAST") == nullptr;  // tests compare both routes
    while (lr.next(&line, &len)) {
        buf.assign(line, line + len);
        buf.push_back('\0');
        unsigned pos = 0, cov = 0;
        const int st = parse_line(buf.data(), len, fast, chrbuf, &pos, &cov, bases, fb.data(), fq.data(), fm.data());
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
