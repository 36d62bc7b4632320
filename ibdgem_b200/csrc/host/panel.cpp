#include "panel.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "textio.h"

namespace ibdhost {

namespace {

std::string zline(const char *line, size_t len) { return std::string(line, len); }

// the reference's "trim off new line char": the LAST character of the line goes, whatever it is
std::string drop_last(const char *line, size_t len) { return len ? std::string(line, len - 1) : std::string(); }

bool is_snp(const char *ref, const char *alt) {  // src/ibdgem.c:113-119
    return strlen(ref) == 1 && strchr("ACGT", ref[0]) && strlen(alt) == 1 && strchr("ACGT", alt[0]);
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

// 16 alleles per step: 32 text bytes "a b c ... " -> compare masks -> the even bits, packed.  Same
// strictness as the scalar paths (every even byte '0' or '1', every odd byte a space); stops at the
// first group that does not fit and returns the number of haplotypes done (a multiple of 16).
#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) static size_t pack_alleles_avx2(const char *hl, size_t hn, size_t H, uint32_t *row) {
    const __m256i c1 = _mm256_set1_epi8('1'), c0 = _mm256_set1_epi8('0'), sp = _mm256_set1_epi8(' ');
    size_t h = 0;
    for (; h + 16 <= H && 2 * h + 32 <= hn; h += 16) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(hl + 2 * h));
        const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, c1));
        const uint32_t m0 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, c0));
        const uint32_t ms = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, sp));
        if (((m0 | m1) & 0x55555555u) != 0x55555555u || (ms & 0xAAAAAAAAu) != 0xAAAAAAAAu) break;
        row[h >> 5] |= _pext_u32(m1, 0x55555555u) << (h & 31);
    }
    return h;
}
// Eight VCF genotype columns of the plain shape "\ta|b" (or "a/b") in 32 bytes -> 16 haplotype bits.
// False if any of the 32 bytes is not what that shape has there (is_gt, src/ibd-parse.c:150-173).
__attribute__((target("avx2,bmi2"))) static bool vcf_group_avx2(const char *p, uint32_t *bits16) {
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('1')));
    const uint32_t m0 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('0')));
    const uint32_t mt = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('\t')));
    const uint32_t ms = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('|'))) |
                        (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('/')));
    if ((mt & 0x11111111u) != 0x11111111u || ((m0 | m1) & 0xAAAAAAAAu) != 0xAAAAAAAAu || (ms & 0x44444444u) != 0x44444444u)
        return false;
    *bits16 = _pext_u32(m1, 0xAAAAAAAAu);
    return true;
}
static const bool kHaveAvx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
#else
static size_t pack_alleles_avx2(const char *, size_t, size_t, uint32_t *) { return 0; }
static bool vcf_group_avx2(const char *, uint32_t *) { return false; }
static const bool kHaveAvx2 = false;
#endif

// One .hap line -> one packed row.  Returns 0, or 1 with the reference-style message in *err.
static int pack_hap_line(const char *hl, size_t hn, size_t H, size_t s, uint32_t *row, std::string *err) {
    char msg[256];
    if (hn < 2 * H - 1) {
        snprintf(msg, sizeof msg, "[::] ERROR: .hap line %zu has %zu characters, %zu haplotypes need %zu.\n", s + 1, hn, H, 2 * H - 1);
        *err = msg;
        return 1;
    }
    size_t h = kHaveAvx2 ? pack_alleles_avx2(hl, hn, H, row) : 0;
    // next: eight text bytes "a b c d " carry four alleles; check the pattern, pick bit 0 of
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
            snprintf(msg, sizeof msg, "[::] ERROR: .hap line %zu: allele '%c' of haplotype %zu is not 0 or 1.\n", s + 1, c, h);
            *err = msg;
            return 1;
        }
    }
    return 0;
}

// "%128s %lu %128s %128s" on a line of the plain shape — four blank-separated tokens, the second one
// to eighteen digits, none longer than 128 characters — without sscanf; false for any other line.
static bool split_legend_fast(const char *p, const char *end, char *id, unsigned long *pos, char *ref, char *alt) {
    auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; };
    if (memchr(p, 0, (size_t)(end - p))) return false;
    const char *tb[4];
    size_t tn[4];
    for (int k = 0; k < 4; k++) {
        while (p < end && blank(*p)) p++;
        const char *q = p;
        while (q < end && !blank(*q)) q++;
        tb[k] = p;
        tn[k] = (size_t)(q - p);
        if (tn[k] < 1 || tn[k] > 128) return false;
        p = q;
    }
    if (tn[1] > 18) return false;
    unsigned long v = 0;
    for (size_t i = 0; i < tn[1]; i++) {
        if (tb[1][i] < '0' || tb[1][i] > '9') return false;
        v = v * 10 + (unsigned long)(tb[1][i] - '0');
    }
    *pos = v;
    memcpy(id, tb[0], tn[0]); id[tn[0]] = 0;
    memcpy(ref, tb[2], tn[2]); ref[tn[2]] = 0;
    memcpy(alt, tb[3], tn[3]); alt[tn[3]] = 0;
    return true;
}

// One legend line -> the per-site text fields (src/ibdgem.c:589-592).
static void push_legend_line(PanelText *out, const char *ll, size_t ln, std::string *lz) {
    char id[129], ref[129], alt[129];
    out->pos.push_back(0);
    out->state.push_back(0);
    out->id_off.push_back(0);
    out->id_len.push_back(0);
    out->ref.push_back('.');
    out->alt.push_back('.');
    unsigned long pos;
    static const bool fast = getenv("IBDGEM_LEGEND_NO_FAST") == nullptr;  // tests compare both routes
    if (fast && split_legend_fast(ll, ll + ln, id, &pos, ref, alt)) {
        // same four values as the sscanf below
    } else {
        lz->assign(ll, ln);
        if (sscanf(lz->c_str(), "%128s %lu %128s %128s", id, &pos, ref, alt) != 4) return;
    }
    out->pos.back() = pos;
    out->state.back() = is_snp(ref, alt) ? 2 : 1;
    out->id_off.back() = out->text.size();
    out->id_len.back() = (uint32_t)strlen(id);
    out->text.append(id, strlen(id) + 1);  // with its NUL
    out->ref.back() = ref[0];
    out->alt.back() = alt[0];
}

// Large plain .hap files: the file is mapped and its lines are packed by several threads (the rows are
// independent; a line's index is the number of newlines before it).  Semantics are those of the
// sequential loop below: lines pair up with legend lines in order, the shorter file ends the panel,
// and the first bad line in file order is the one reported.
static constexpr size_t MT_MIN_BYTES = (size_t)32 << 20;
static constexpr unsigned MT_MAX_THREADS = 16;
static size_t mt_min_bytes() {  // IBDGEM_PACK_MT_MIN_BYTES: tests push small fixtures through the threaded path
    static size_t v = 0;
    if (!v) {
        const char *sm = getenv("IBDGEM_PACK_MT_MIN_BYTES");
        v = sm && atol(sm) > 0 ? (size_t)atol(sm) : MT_MIN_BYTES;
    }
    return v;
}

static int parse_hap_mapped(const std::string &hap_fn, size_t H, PanelText *out, size_t n_legend, bool *done) {
    *done = false;
    const int fd = open(hap_fn.c_str(), O_RDONLY);
    if (fd < 0) return 0;  // the caller's sequential path reports it
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < mt_min_bytes() || st.st_size == 0) {
        close(fd);
        return 0;
    }
    const size_t n = (size_t)st.st_size;
    void *map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return 0;
    madvise(map, n, MADV_SEQUENTIAL);
    const char *data = static_cast<const char *>(map);
    const unsigned nt = std::max(1u, std::min(MT_MAX_THREADS, std::thread::hardware_concurrency()));
    std::vector<size_t> cut(nt + 1), newlines(nt, 0);
    for (unsigned t = 0; t <= nt; t++) cut[t] = n / nt * t;
    cut[nt] = n;
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                size_t c = 0;
                for (const char *p = data + cut[t], *e = data + cut[t + 1]; p < e;) {
                    const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(e - p)));
                    if (!nl) break;
                    c++;
                    p = nl + 1;
                }
                newlines[t] = c;
            });
        for (auto &x : th) x.join();
    }
    // first_line[t] = newlines before cut[t]: the index of the line that contains byte cut[t]
    size_t n_lines = 0;
    std::vector<size_t> first_line(nt);
    for (unsigned t = 0; t < nt; t++) {
        first_line[t] = n_lines;
        n_lines += newlines[t];
    }
    if (n > 0 && data[n - 1] != '\n') n_lines++;  // last line without a newline
    const size_t S = std::min(n_lines, n_legend);
    out->S = (int64_t)S;
    out->bits.assign(S * (size_t)out->Wh, 0u);
    std::vector<std::string> errs(nt);
    std::vector<size_t> err_line(nt, ~(size_t)0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                size_t p = cut[t], idx = first_line[t];
                if (t > 0 && data[p - 1] != '\n') {  // move to the first line that starts in this range
                    const char *nl = static_cast<const char *>(memchr(data + p, '\n', cut[t + 1] - p));
                    if (!nl) return;  // the range lies inside one line
                    p = (size_t)(nl - data) + 1;
                    idx++;
                }
                while (p < cut[t + 1] && idx < S) {
                    const char *nl = static_cast<const char *>(memchr(data + p, '\n', n - p));
                    const size_t hn = nl ? (size_t)(nl - (data + p)) + 1 : n - p;
                    if (pack_hap_line(data + p, hn, H, idx, out->bits.data() + idx * (size_t)out->Wh, &errs[t])) {
                        err_line[t] = idx;
                        return;
                    }
                    p += hn;
                    idx++;
                }
            });
        for (auto &x : th) x.join();
    }
    munmap(map, n);
    for (unsigned t = 0; t < nt; t++)  // ranges are in file order: the first recorded error is the first bad line
        if (err_line[t] != ~(size_t)0) {
            fputs(errs[t].c_str(), stderr);
            return 1;
        }
    *done = true;
    return 0;
}

int parse_impute(const std::string &hap_fn, const std::string &legend_fn, const std::vector<std::string> &names,
                 PanelText *out) {
    LineReader hap, leg;
    if (!hap.open(hap_fn) || !leg.open(legend_fn)) {
        fprintf(stderr, "[::] ERROR parsing hap/legend/indv data; make sure inputs are valid.\n");
        return 1;
    }
    out->names = names;
    out->N = (int32_t)names.size();
    const int64_t words = (2 * (int64_t)out->N + 31) / 32;
    out->Wh = (words + 3) / 4 * 4;  // 16-byte rows for 128-bit loads on the device
    const size_t H = 2 * names.size();
    const char *hl, *ll;
    size_t hn, ln;
    leg.next(&ll, &ln);  // header (src/ibdgem.c:554)
    std::string lz, err;
    if (!ends_with_gz(hap_fn) && H > 0) {
        // the legend is small: read all of it, then pack the mapped .hap lines in parallel
        while (leg.next(&ll, &ln)) push_legend_line(out, ll, ln, &lz);
        bool done = false;
        if (parse_hap_mapped(hap_fn, H, out, out->pos.size(), &done)) return 1;
        if (done) {
            const size_t S = (size_t)out->S;  // the shorter file ends the panel
            out->pos.resize(S); out->state.resize(S); out->id_off.resize(S); out->id_len.resize(S);
            out->ref.resize(S); out->alt.resize(S);
            return 0;
        }
        // small file: the sequential loop, over the legend lines already read
        const size_t n_leg = out->pos.size();
        size_t s = 0;
        for (; s < n_leg && hap.next(&hl, &hn); s++) {
            out->bits.resize(out->bits.size() + (size_t)out->Wh, 0u);
            out->S++;
            if (pack_hap_line(hl, hn, H, s, out->bits.data() + s * (size_t)out->Wh, &err)) {
                fputs(err.c_str(), stderr);
                return 1;
            }
        }
        out->pos.resize(s); out->state.resize(s); out->id_off.resize(s); out->id_len.resize(s);
        out->ref.resize(s); out->alt.resize(s);
        return 0;
    }
    while (hap.next(&hl, &hn) && leg.next(&ll, &ln)) {
        const size_t s = (size_t)out->S;
        out->bits.resize(out->bits.size() + (size_t)out->Wh, 0u);
        out->S++;
        if (pack_hap_line(hl, hn, H, s, out->bits.data() + s * (size_t)out->Wh, &err)) {
            fputs(err.c_str(), stderr);
            return 1;
        }
        push_legend_line(out, ll, ln, &lz);
    }
    return 0;
}

int join_pileup(PanelText *pt, const PileupStore &pu, const PackOptions &opt, PackedPanel *out) {
    const size_t S = (size_t)pt->S;
    out->S = pt->S;
    out->N = pt->N;
    out->Wh = pt->Wh;
    out->names = pt->names;
    out->pos = pt->pos;
    out->bits = std::move(pt->bits);
    out->n_ref.assign(S, 0);
    out->n_alt.assign(S, 0);
    out->host_keep.assign(S, 0);
    out->dp.assign(S, 0);
    out->chr_id.assign(S, 0);
    out->id_off.assign(S, 0);
    out->id_len.assign(S, 0);
    out->ref.assign(S, '.');
    out->alt.assign(S, '.');
    for (size_t s = 0; s < S; s++) {
        if (pt->state[s] != 2) continue;  // is_snp, src/ibdgem.c:592
        if (!pt->qual.empty() && pt->qual[s] < opt.min_qual) continue;  // VCF -q, src/ibdgem.c:297
        const uint64_t pos = pt->pos[s];
        const int64_t pul = pu.fetch(pos);
        if (pul < 0) continue;
        if (opt.positions && !opt.positions->count(pos)) continue;
        fill_kept(out, s, pu, pul, pt->text.c_str() + pt->id_off[s], pt->ref[s], pt->alt[s]);
    }
    if (opt.af) {
        out->af_user.assign(S, NAN);
        for (size_t s = 0; s < S; s++)
            if (out->host_keep[s])
                if (const double *f = opt.af->fetch(out->pos[s])) out->af_user[s] = *f;
    }
    return 0;
}

int pack_impute(const std::string &hap_fn, const std::string &legend_fn, const std::vector<std::string> &names,
                const PileupStore &pu, const PackOptions &opt, PackedPanel *out) {
    PanelText pt;
    if (parse_impute(hap_fn, legend_fn, names, &pt)) return 1;
    return join_pileup(&pt, pu, opt, out);
}

// --- binary cache ------------------------------------------------------------------------------
namespace {

struct CacheHeader {
    char magic[8];
    uint64_t key[6];  // size and mtime (ns) of .hap, .legend, .indv
    int64_t S;
    int64_t Wh;
    int32_t N;
    uint32_t has_qual;  // VCF panels carry the QUAL column (the -q filter is applied at join time)
    uint64_t names_bytes, text_bytes;
};
const char kCacheMagic[8] = {'I', 'B', 'D', 'G', 'P', 'N', 'L', '2'};

bool file_key(const std::string &fn, uint64_t *size, uint64_t *mtime_ns) {
    struct stat st;
    if (stat(fn.c_str(), &st) != 0) return false;
    *size = (uint64_t)st.st_size;
    *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    return true;
}
// up to three input files (.hap, .legend, .indv — or the VCF alone, the other slots zero)
bool make_key(const std::vector<std::string> &inputs, uint64_t key[6]) {
    memset(key, 0, 6 * sizeof(uint64_t));
    if (inputs.empty() || inputs.size() > 3) return false;
    for (size_t i = 0; i < inputs.size(); i++)
        if (!file_key(inputs[i], &key[2 * i], &key[2 * i + 1])) return false;
    return true;
}
template <class T>
bool put(FILE *f, const std::vector<T> &v) { return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size(); }
template <class T>
bool get(FILE *f, std::vector<T> *v, size_t n) {
    v->resize(n);
    return n == 0 || fread(v->data(), sizeof(T), n, f) == n;
}

}  // namespace

int save_panel_cache(const std::string &cache_fn, const std::vector<std::string> &inputs, const PanelText &pt) {
    CacheHeader h{};
    memcpy(h.magic, kCacheMagic, 8);
    if (!make_key(inputs, h.key)) return 1;
    h.has_qual = pt.qual.empty() ? 0u : 1u;
    std::string names;
    for (const auto &n : pt.names) names.append(n.c_str(), n.size() + 1);
    h.S = pt.S; h.Wh = pt.Wh; h.N = pt.N;
    h.names_bytes = names.size();
    h.text_bytes = pt.text.size();
    const std::string tmp = cache_fn + ".tmp";  // a reader never sees a half-written cache
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return 1;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && (names.empty() || fwrite(names.data(), 1, names.size(), f) == names.size()) &&
              put(f, pt.pos) && put(f, pt.state) && put(f, pt.id_off) && put(f, pt.id_len) && put(f, pt.ref) && put(f, pt.alt) &&
              put(f, pt.qual) &&
              (pt.text.empty() || fwrite(pt.text.data(), 1, pt.text.size(), f) == pt.text.size()) && put(f, pt.bits);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), cache_fn.c_str()) != 0) {
        remove(tmp.c_str());
        return 1;
    }
    return 0;
}

bool load_panel_cache(const std::string &cache_fn, const std::vector<std::string> &inputs, PanelText *pt) {
    FILE *f = fopen(cache_fn.c_str(), "rb");
    if (!f) return false;
    CacheHeader h;
    uint64_t key[6];
    bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, kCacheMagic, 8) == 0 &&
              make_key(inputs, key) && memcmp(key, h.key, sizeof key) == 0 && h.S >= 0 && h.N > 0 &&
              h.Wh * 32 >= 2 * (int64_t)h.N;
    if (ok) {
        const size_t S = (size_t)h.S;
        std::vector<char> names, text;
        ok = get(f, &names, (size_t)h.names_bytes) && get(f, &pt->pos, S) && get(f, &pt->state, S) && get(f, &pt->id_off, S) &&
             get(f, &pt->id_len, S) && get(f, &pt->ref, S) && get(f, &pt->alt, S) && get(f, &pt->qual, h.has_qual ? S : 0) &&
             get(f, &text, (size_t)h.text_bytes) &&
             get(f, &pt->bits, S * (size_t)h.Wh) && fgetc(f) == EOF;
        if (ok) {
            pt->S = h.S; pt->N = h.N; pt->Wh = h.Wh;
            pt->text.assign(text.data(), text.size());
            pt->names.clear();
            for (size_t i = 0; i < names.size();) {
                const size_t n = strnlen(names.data() + i, names.size() - i);
                pt->names.emplace_back(names.data() + i, n);
                i += n + 1;
            }
            ok = (int64_t)pt->names.size() == (int64_t)h.N;
            for (size_t s = 0; ok && s < S; s++)  // offsets must stay inside the text blob
                ok = pt->state[s] == 0 || (pt->id_off[s] + pt->id_len[s] < pt->text.size() && pt->text[pt->id_off[s] + pt->id_len[s]] == 0);
        }
    }
    fclose(f);
    if (!ok) *pt = PanelText();
    return ok;
}

int pack_impute_cached(const std::string &hap_fn, const std::string &legend_fn, const std::string &indv_fn,
                       const std::string &cache_fn, const PileupStore &pu, const PackOptions &opt, PackedPanel *out,
                       bool *hit) {
    const bool timing = getenv("IBDGEM_PACK_TIMING") != nullptr;  // stage times on stderr
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    PanelText pt;
    const std::vector<std::string> inputs{hap_fn, legend_fn, indv_fn};
    const bool cached = load_panel_cache(cache_fn, inputs, &pt);
    if (hit) *hit = cached;
    const double t1 = now();
    double t2 = t1;
    if (!cached) {
        std::vector<std::string> names;
        if (read_indv(indv_fn, &names) || parse_impute(hap_fn, legend_fn, names, &pt)) return 1;
        t2 = now();
        if (save_panel_cache(cache_fn, inputs, pt))
            fprintf(stderr, "[::] WARNING: could not write the panel cache %s.\n", cache_fn.c_str());
    }
    const double t3 = now();
    const int rc = join_pileup(&pt, pu, opt, out);
    if (timing)
        fprintf(stderr, "[pack] cache load %.3f s, text parse %.3f s, cache write %.3f s, pileup join %.3f s\n", t1 - t0, t2 - t1,
                t3 - t2, now() - t3);
    return rc;
}

// --- VCF ------------------------------------------------------------------------------------------
namespace {

// What one VCF record contributes to the panel text (the bits go straight into the record's row).
struct VcfRecord {
    uint64_t pos = 0;
    uint8_t state = 0;
    char ref = '.', alt = '.';
    double qual = 0;
    size_t id_off = 0;  // into the text blob the id was appended to
    uint32_t id_len = 0;
};

struct VcfScratch {
    std::string id, ref, alt, qual;
};

// One body line (src/ibdgem.c:272-331).  `text` receives the record's ID with a NUL when the record
// reaches the SNP test; `msgs` the reference's diagnostics.
void parse_vcf_line(const char *line, size_t len, size_t N, int64_t Wh, uint32_t *row, VcfRecord *rec, std::string *text,
                    std::string *msgs, VcfScratch *sc) {
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
    if (nf < 9 || i > e) return;  // fewer than 10 columns: "skipped"
    char *endp = nullptr;
    const std::string posz(line + fb[1], fe[1] - fb[1]);
    const unsigned long pos = strtoul(posz.c_str(), &endp, 10);
    if (endp == posz.c_str()) return;
    rec->pos = pos;
    std::string &id = sc->id, &ref = sc->ref, &alt = sc->alt, &qual = sc->qual;
    id.assign(line + fb[2], fe[2] - fb[2]);
    ref.assign(line + fb[3], fe[3] - fb[3]);
    alt.assign(line + fb[4], fe[4] - fb[4]);
    qual.assign(line + fb[5], fe[5] - fb[5]);
    if (id.empty() || ref.empty() || alt.empty() || qual.empty() || fe[6] == fb[6] || fe[7] == fb[7]) return;
    if (alt.find(',') != std::string::npos) return;  // is_biallelic, src/ibdgem.c:161-167
    // genotypes of EVERY sample must look like [01][/|][01].* (src/ibd-parse.c:150-173)
    size_t k = 0;
    bool ok = true;
    while (k < N) {
        if (kHaveAvx2 && k + 8 <= N) {
            // eight plain "\ta|b" columns at once; anything else (longer fields, doubled tabs, bad
            // characters) is left to the column-by-column code below
            const size_t p = (i < e && line[i] == '\t') ? i : i - 1;  // (i > 0 here: nine columns came before)
            uint32_t got;
            if (line[p] == '\t' && p + 32 <= e && vcf_group_avx2(line + p, &got)) {
                const uint64_t v = (uint64_t)got << ((2 * k) & 31);
                row[(2 * k) >> 5] |= (uint32_t)v;
                if (v >> 32) row[((2 * k) >> 5) + 1] |= (uint32_t)(v >> 32);
                k += 8;
                i = p + 32;
                while (i < e && line[i] != '\t') i++;  // the eighth column may carry more than the genotype
                continue;
            }
        }
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
        char m[128];
        snprintf(m, sizeof m, "Failed to parse genotype fields at %lu. Skipping to next site.\n", pos);
        msgs->append(m);
        for (int64_t w = 0; w < Wh; w++) row[w] = 0;
        return;
    }
    // what is left depends on the options and the pileup (join_pileup): SNP test, -q, pileup line, -p
    rec->state = is_snp(ref.c_str(), alt.c_str()) ? 2 : 1;
    rec->qual = atof(qual.c_str());
    rec->id_off = text->size();
    rec->id_len = (uint32_t)id.size();
    text->append(id.c_str(), id.size() + 1);  // with its NUL
    rec->ref = ref[0];
    rec->alt = alt[0];
}

void store_record(PanelText *out, size_t s, const VcfRecord &r, size_t text_base) {
    out->pos[s] = r.pos;
    out->state[s] = r.state;
    out->ref[s] = r.ref;
    out->alt[s] = r.alt;
    out->qual[s] = r.qual;
    out->id_off[s] = r.state ? text_base + r.id_off : 0;
    out->id_len[s] = r.id_len;
}

void size_records(PanelText *out, size_t S) {
    out->S = (int64_t)S;
    out->pos.assign(S, 0);
    out->state.assign(S, 0);
    out->id_off.assign(S, 0);
    out->id_len.assign(S, 0);
    out->ref.assign(S, '.');
    out->alt.assign(S, '.');
    out->qual.assign(S, 0.0);
    out->bits.assign(S * (size_t)out->Wh, 0u);
}

// sample names of the "#CHROM" line, split on tabs like strtok (src/ibd-parse.c:113-147)
int parse_vcf_header(const char *line, size_t len, PanelText *out) {
    static const char kHead[] = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t";
    const size_t hl = sizeof(kHead) - 1;
    if (!line || len <= hl || memcmp(line, kHead, hl) != 0 || line[hl] == '\n') {
        fprintf(stderr, "[::] ERROR parsing VCF header.\n");
        return 1;
    }
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
    out->N = (int32_t)out->names.size();
    out->Wh = ((2 * (int64_t)out->N + 31) / 32 + 3) / 4 * 4;  // 16-byte rows for 128-bit loads on the device
    return 0;
}

// Large plain VCF files: mapped, the body lines parsed by several threads (records are independent; a
// record's index is the number of newlines between the header and it).  IDs and diagnostics are
// gathered per thread and joined in file order, so the result is that of the sequential reader.
int parse_vcf_mapped(const std::string &vcf_fn, PanelText *out, bool *done) {
    *done = false;
    const int fd = open(vcf_fn.c_str(), O_RDONLY);
    if (fd < 0) return 0;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < mt_min_bytes() || st.st_size == 0) {
        close(fd);
        return 0;
    }
    const size_t n = (size_t)st.st_size;
    void *map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return 0;
    madvise(map, n, MADV_SEQUENTIAL);
    const char *data = static_cast<const char *>(map);
    // header: "##" lines, then the "#CHROM" line
    size_t p = 0;
    const char *hline = nullptr;
    size_t hlen = 0;
    while (p < n) {
        const char *nl = static_cast<const char *>(memchr(data + p, '\n', n - p));
        const size_t len = nl ? (size_t)(nl - (data + p)) + 1 : n - p;
        if (len >= 2 && data[p] == '#' && data[p + 1] == '#') {
            p += len;
            continue;
        }
        hline = data + p;
        hlen = len;
        p += len;
        break;
    }
    if (parse_vcf_header(hline, hlen, out)) {
        munmap(map, n);
        return 1;
    }
    const size_t body = p, N = out->names.size();
    const unsigned nt = std::max(1u, std::min(MT_MAX_THREADS, std::thread::hardware_concurrency()));
    std::vector<size_t> cut(nt + 1), newlines(nt, 0);
    for (unsigned t = 0; t <= nt; t++) cut[t] = body + (n - body) / nt * t;
    cut[nt] = n;
    auto for_threads = [&](auto fn) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    for_threads([&](unsigned t) {
        size_t c = 0;
        for (const char *q = data + cut[t], *e = data + cut[t + 1]; q < e;) {
            const char *nl = static_cast<const char *>(memchr(q, '\n', (size_t)(e - q)));
            if (!nl) break;
            c++;
            q = nl + 1;
        }
        newlines[t] = c;
    });
    size_t S = 0;
    std::vector<size_t> first_line(nt);  // newlines of the body before cut[t]
    for (unsigned t = 0; t < nt; t++) {
        first_line[t] = S;
        S += newlines[t];
    }
    if (n > body && data[n - 1] != '\n') S++;  // last record without a newline
    size_records(out, S);
    struct Part {
        std::string text, msgs;
        std::vector<VcfRecord> recs;
        size_t first = 0;
    };
    std::vector<Part> parts(nt);
    for_threads([&](unsigned t) {
        Part &pt = parts[t];
        size_t q = cut[t], idx = first_line[t];
        if (q > body && data[q - 1] != '\n') {  // move to the first line that starts in this range
            const char *nl = static_cast<const char *>(memchr(data + q, '\n', cut[t + 1] - q));
            if (!nl) return;
            q = (size_t)(nl - data) + 1;
            idx++;
        }
        pt.first = idx;
        VcfScratch sc;
        while (q < cut[t + 1] && idx < S) {
            const char *nl = static_cast<const char *>(memchr(data + q, '\n', n - q));
            const size_t len = nl ? (size_t)(nl - (data + q)) + 1 : n - q;
            pt.recs.emplace_back();
            parse_vcf_line(data + q, len, N, out->Wh, out->bits.data() + idx * (size_t)out->Wh, &pt.recs.back(), &pt.text, &pt.msgs, &sc);
            q += len;
            idx++;
        }
    });
    munmap(map, n);
    for (unsigned t = 0; t < nt; t++) {
        const size_t base = out->text.size();
        out->text += parts[t].text;
        for (size_t k = 0; k < parts[t].recs.size(); k++) store_record(out, parts[t].first + k, parts[t].recs[k], base);
        fputs(parts[t].msgs.c_str(), stderr);
    }
    *done = true;
    return 0;
}

}  // namespace

int parse_vcf(const std::string &vcf_fn, PanelText *out) {
    if (!ends_with_gz(vcf_fn)) {
        {   // same diagnostics as the sequential reader for a file that cannot be opened
            LineReader probe;
            if (!probe.open(vcf_fn)) return 1;
        }
        bool done = false;
        if (parse_vcf_mapped(vcf_fn, out, &done)) return 1;
        if (done) return 0;
        *out = PanelText();
    }
    LineReader vcf;
    if (!vcf.open(vcf_fn)) return 1;
    const char *line;
    size_t len;
    bool have = vcf.next(&line, &len);
    while (have && len >= 2 && line[0] == '#' && line[1] == '#') have = vcf.next(&line, &len);
    if (parse_vcf_header(have ? line : nullptr, have ? len : 0, out)) return 1;
    const size_t N = out->names.size();
    VcfScratch sc;
    std::string msgs;
    while (vcf.next(&line, &len)) {
        const size_t s = (size_t)out->S;
        out->pos.push_back(0);
        out->state.push_back(0);
        out->id_off.push_back(0);
        out->id_len.push_back(0);
        out->ref.push_back('.');
        out->alt.push_back('.');
        out->qual.push_back(0.0);
        out->bits.resize(out->bits.size() + (size_t)out->Wh, 0u);
        out->S++;
        VcfRecord rec;
        parse_vcf_line(line, len, N, out->Wh, out->bits.data() + s * (size_t)out->Wh, &rec, &out->text, &msgs, &sc);
        store_record(out, s, rec, 0);
        if (!msgs.empty()) {
            fputs(msgs.c_str(), stderr);
            msgs.clear();
        }
    }
    return 0;
}

int pack_vcf(const std::string &vcf_fn, const PileupStore &pu, const PackOptions &opt, PackedPanel *out) {
    PanelText pt;
    if (parse_vcf(vcf_fn, &pt)) return 1;
    return join_pileup(&pt, pu, opt, out);
}

int pack_vcf_cached(const std::string &vcf_fn, const std::string &cache_fn, const PileupStore &pu, const PackOptions &opt,
                    PackedPanel *out, bool *hit) {
    PanelText pt;
    const std::vector<std::string> inputs{vcf_fn};
    const bool cached = load_panel_cache(cache_fn, inputs, &pt);
    if (hit) *hit = cached;
    if (!cached) {
        if (parse_vcf(vcf_fn, &pt)) return 1;
        if (save_panel_cache(cache_fn, inputs, pt))
            fprintf(stderr, "[::] WARNING: could not write the panel cache %s.\n", cache_fn.c_str());
    }
    return join_pileup(&pt, pu, opt, out);
}

}  // namespace ibdhost
