// host_abi.cpp — C entry points over the host packer (no CUDA), so the CPU test-suite can check
// the parsers and the bit packing against the reference's fixtures without a GPU.
#include <cstdio>
#include <cstring>

#include "panel.h"
#include "pileup_store.h"

using namespace ibdhost;

namespace {
struct Handle {
    PileupStore pu;
    PackedPanel panel;
    FreqTable af;
    std::unordered_set<uint64_t> positions;
};
}  // namespace

extern "C" {

// mode 0 = IMPUTE (a = .hap, b = .legend, c = .indv), 1 = VCF (a = file).  NULL on any error.
void *ibdhost_pack(int mode, const char *a, const char *b, const char *c, const char *pileup, const char *chr,
                   const char *positions_fn, const char *af_fn, double min_qual) {
    Handle *h = new Handle();
    PackOptions po;
    po.min_qual = min_qual;
    bool ok = load_pileup(pileup, chr, &h->pu) == 0;
    if (ok && af_fn) {
        ok = read_af(af_fn, chr, &h->af) == 0;
        po.af = &h->af;
    }
    if (ok && positions_fn) {
        ok = read_positions(positions_fn, chr, &h->positions) == 0;
        po.positions = &h->positions;
    }
    if (ok && mode == 0) {
        std::vector<std::string> names;
        ok = read_indv(c, &names) == 0 && pack_impute(a, b, names, h->pu, po, &h->panel) == 0;
    } else if (ok) {
        ok = pack_vcf(a, h->pu, po, &h->panel) == 0;
    }
    if (!ok) {
        delete h;
        return nullptr;
    }
    return h;
}
// IMPUTE panel through the binary cache; *hit = 1 when the text files were not parsed.
void *ibdhost_pack_cached(const char *hap, const char *legend, const char *indv, const char *cache, const char *pileup,
                          const char *chr, int *hit) {
    Handle *h = new Handle();
    PackOptions po;
    bool was_hit = false;
    const bool ok = load_pileup(pileup, chr, &h->pu) == 0 &&
                    pack_impute_cached(hap, legend, indv, cache, h->pu, po, &h->panel, &was_hit) == 0;
    if (hit) *hit = was_hit ? 1 : 0;
    if (!ok) {
        delete h;
        return nullptr;
    }
    return h;
}
void *ibdhost_pack_vcf_cached(const char *vcf, const char *cache, const char *pileup, const char *chr, double min_qual,
                              int *hit) {
    Handle *h = new Handle();
    PackOptions po;
    po.min_qual = min_qual;
    bool was_hit = false;
    const bool ok = load_pileup(pileup, chr, &h->pu) == 0 && pack_vcf_cached(vcf, cache, h->pu, po, &h->panel, &was_hit) == 0;
    if (hit) *hit = was_hit ? 1 : 0;
    if (!ok) {
        delete h;
        return nullptr;
    }
    return h;
}
void ibdhost_free(void *p) { delete static_cast<Handle *>(p); }
int64_t ibdhost_n_sites(void *p) { return static_cast<Handle *>(p)->panel.S; }
int32_t ibdhost_n_indiv(void *p) { return static_cast<Handle *>(p)->panel.N; }
int64_t ibdhost_words(void *p) { return static_cast<Handle *>(p)->panel.Wh; }
int64_t ibdhost_n_pileup(void *p) { return (int64_t)static_cast<Handle *>(p)->pu.size(); }
const uint64_t *ibdhost_pos(void *p) { return static_cast<Handle *>(p)->panel.pos.data(); }
const uint8_t *ibdhost_n_ref(void *p) { return static_cast<Handle *>(p)->panel.n_ref.data(); }
const uint8_t *ibdhost_n_alt(void *p) { return static_cast<Handle *>(p)->panel.n_alt.data(); }
const uint8_t *ibdhost_keep(void *p) { return static_cast<Handle *>(p)->panel.host_keep.data(); }
const uint32_t *ibdhost_dp(void *p) { return static_cast<Handle *>(p)->panel.dp.data(); }
const uint32_t *ibdhost_bits(void *p) { return static_cast<Handle *>(p)->panel.bits.data(); }
const double *ibdhost_af_user(void *p) {
    Handle *h = static_cast<Handle *>(p);
    return h->panel.af_user.empty() ? nullptr : h->panel.af_user.data();
}
const uint32_t *ibdhost_pileup_cov(void *p) { return static_cast<Handle *>(p)->pu.cov.data(); }
// "chr\trsID\tREF\tALT" of a kept site, as the tab.txt writer would print them; 0 for other sites
int ibdhost_site_label(void *p, int64_t site, char *buf, int cap) {
    Handle *h = static_cast<Handle *>(p);
    const PackedPanel &P = h->panel;
    const size_t s = (size_t)site;
    if (site < 0 || site >= P.S || !P.host_keep[s]) return 0;
    return snprintf(buf, (size_t)cap, "%s\t%.*s\t%c\t%c", h->pu.chr_names[P.chr_id[s]].c_str(), (int)P.id_len[s],
                    P.text.data() + P.id_off[s], P.ref[s], P.alt[s]);
}
const char *ibdhost_name(void *p, int32_t i) { return static_cast<Handle *>(p)->panel.names[(size_t)i].c_str(); }

}  // extern "C"
